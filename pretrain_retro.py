"""Pretrain Retro (reference ``pretrain_retro.py``): GPT decoder with chunked cross-attention over
retrieved neighbours.  Neighbour tokens come from the dataset sample (``neighbor_tokens``); without a
retrieval database (``tools/retro`` in the reference builds one with faiss) mock neighbours are used."""
import os

if not os.path.exists("/dev/nvidia0"):
    os.environ.setdefault("OMP_NUM_THREADS", "2")

from functools import partial

import torch

from megatronapp_b200.models.retro import RetroConfig, RetroModel
from megatronapp_b200.parallel import state as mpu
from megatronapp_b200.training.arguments import core_transformer_config_from_args
from megatronapp_b200.training.global_vars import get_args
from megatronapp_b200.training.training import pretrain
from megatronapp_b200.training.utils import print_rank_0


def add_retro_args(parser):
    g = parser.add_argument_group("retro")
    g.add_argument("--retro-chunk-length", type=int, default=64)
    g.add_argument("--retro-num-neighbors", type=int, default=2)
    g.add_argument("--retro-retrieved-length", type=int, default=128)
    g.add_argument("--retro-encoder-layers", type=int, default=2)
    g.add_argument("--retro-num-retrieved-chunks", type=int, default=None,
                   help="retrieved length in chunks (neighbour + continuation = 2); overrides --retro-retrieved-length")
    g.add_argument("--retro-attention-gate", type=float, default=1.0)
    g.add_argument("--retro-no-verify-neighbor-count", action="store_false", dest="retro_verify_neighbor_count")
    return parser


def _check_parallelism(args):
    if args.pipeline_model_parallel_size != 1:
        raise ValueError("retro currently does not support pipeline parallelism (same restriction as the reference)")


def model_provider(pre_process=True, post_process=True):
    args = get_args()
    _check_parallelism(args)
    # --retro-add-retriever is implied: this entry point always builds the retriever (the reference falls back to a plain GPT
    # without the flag; that baseline is pretrain_gpt.py here).  With the flag a GPT checkpoint loads non-strictly.
    print_rank_0("building Retro model ...")
    if args.retro_num_retrieved_chunks:
        args.retro_retrieved_length = args.retro_num_retrieved_chunks * args.retro_chunk_length
    config = core_transformer_config_from_args(args, RetroConfig)
    config.retro_attention_gate, config.retro_verify_neighbor_count = args.retro_attention_gate, args.retro_verify_neighbor_count
    config.retro_chunk_length, config.retro_num_neighbors = args.retro_chunk_length, args.retro_num_neighbors
    config.retro_retrieved_length, config.retro_encoder_num_layers = args.retro_retrieved_length, args.retro_encoder_layers
    return RetroModel(config, vocab_size=args.padded_vocab_size, max_sequence_length=args.max_position_embeddings,
                      pre_process=pre_process, post_process=post_process, parallel_output=True,
                      share_embeddings_and_output_weights=not args.untie_embeddings_and_output_weights,
                      position_embedding_type=args.position_embedding_type)


class MockRetroDataset(torch.utils.data.Dataset):
    def __init__(self, n, seq, vocab, chunk, k, r, seed):
        self.n, self.seq, self.vocab, self.chunk, self.k, self.r, self.seed = n, seq, vocab, chunk, k, r, seed

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        g = torch.Generator().manual_seed(self.seed + i)
        toks = torch.randint(1, self.vocab, (self.seq + 1,), generator=g)
        nb = torch.randint(1, self.vocab, (self.seq // self.chunk, self.k, self.r), generator=g)
        return {"tokens": toks[:-1].clone(), "labels": toks[1:].clone(), "loss_mask": torch.ones(self.seq),
                "position_ids": torch.arange(self.seq), "neighbor_tokens": nb}


def train_valid_test_datasets_provider(num_samples):
    a = get_args()
    if a.retro_project_dir and a.data_path:
        # real data: the GPT blend + the neighbour table tools/retro/preprocess_data.py wrote for exactly this blend / split /
        # seed / sequence length; validation and test fall back to mock neighbours (only the train split is queried)
        from megatronapp_b200.data.gpt_dataset import build_gpt_datasets
        from megatronapp_b200.data.retro import ChunkDB, RetroDataset, RetroPreprocessingConfig

        cfg = RetroPreprocessingConfig.load(a.retro_project_dir)
        for mine, theirs, what in ((a.seq_length, cfg.retro_gpt_seq_length, "sequence length"),
                                   (a.retro_chunk_length, cfg.retro_gpt_chunk_length, "chunk length"),
                                   (a.retro_retrieved_length, cfg.retro_retrieved_length, "retrieved length"),
                                   (a.retro_num_neighbors, cfg.retro_query_num_neighbors_save, "neighbour count")):
            if mine != theirs:
                raise ValueError(f"retro project was preprocessed with {what} {theirs}, training asks for {mine}")
        train, valid, test = build_gpt_datasets(a, [cfg.retro_gpt_train_samples or num_samples[0], num_samples[1], num_samples[2]])
        mock = lambda n, s: MockRetroDataset(max(int(n), 1), a.seq_length, a.padded_vocab_size, a.retro_chunk_length,  # noqa: E731
                                             a.retro_num_neighbors, a.retro_retrieved_length, s)
        return RetroDataset(cfg, train, ChunkDB(cfg), "train", verify_neighbor_count=a.retro_verify_neighbor_count), mock(num_samples[1], 2), mock(num_samples[2], 3)
    mk = lambda n, s: MockRetroDataset(max(int(n), 1), a.seq_length, a.padded_vocab_size, a.retro_chunk_length,  # noqa: E731
                                       a.retro_num_neighbors, a.retro_retrieved_length, s)
    return mk(num_samples[0], 1), mk(num_samples[1], 2), mk(num_samples[2], 3)


def loss_func(loss_mask, output_tensor):
    losses = output_tensor.float()
    loss_mask = loss_mask.reshape(-1).float()
    loss = torch.cat([torch.sum(losses.reshape(-1) * loss_mask).view(1), loss_mask.sum().view(1)])
    rep = loss.clone().detach()
    g = mpu.get_data_parallel_group()
    if g is not None and not isinstance(g, mpu._SingleGroup) and torch.distributed.get_world_size(g) > 1:
        torch.distributed.all_reduce(rep, group=g)
    return loss[0], loss[1].clone().detach().to(torch.int), {"lm loss": (rep[0], rep[1])}


def forward_step(data_iterator, model):
    dev = next(model.parameters()).device
    d = next(data_iterator)
    out = model(d["tokens"].to(dev), d["position_ids"].to(dev), None, context_input_ids=d["neighbor_tokens"].to(dev),
                labels=d["labels"].to(dev))
    return out, partial(loss_func, d["loss_mask"].to(dev))


if __name__ == "__main__":
    train_valid_test_datasets_provider.is_distributed = True
    pretrain(train_valid_test_datasets_provider, model_provider, "encoder_or_decoder", forward_step,
             extra_args_provider=add_retro_args, args_defaults={"tokenizer_type": "NullTokenizer"})
