"""Pretrain a BERT dual encoder with the inverse cloze task (entry point kept from the reference ``pretrain_ict.py``)."""
import os

if not os.path.exists("/dev/nvidia0"):
    os.environ.setdefault("OMP_NUM_THREADS", "2")

from functools import partial

import numpy as np
import torch

from megatronapp_b200.legacy.model.biencoder_model import BiEncoderModel, ict_loss
from megatronapp_b200.parallel import state as mpu
from megatronapp_b200.training.arguments import core_transformer_config_from_args
from megatronapp_b200.training.global_vars import get_args, get_tokenizer
from megatronapp_b200.training.training import pretrain
from megatronapp_b200.training.utils import print_rank_0


class MockICTDataset(torch.utils.data.Dataset):
    """Synthetic retrieval pairs: query and context of sample ``i`` draw tokens from the same small vocabulary slice."""

    def __init__(self, n, seq, vocab, seed):
        self.n, self.seq, self.vocab, self.seed = n, seq, vocab, seed

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        rng = np.random.default_rng(self.seed + i)
        lo = 5 + (i * 7) % max(1, self.vocab - 21)
        q = np.concatenate([[1], rng.integers(lo, lo + 16, self.seq // 4), [2]])
        c = np.concatenate([[1], rng.integers(lo, lo + 16, self.seq - 2), [2]])
        pad = lambda a: (np.pad(a, (0, self.seq - len(a))), (np.arange(self.seq) < len(a)).astype(np.int64))  # noqa: E731
        (qt, qm), (ct, cm) = pad(q), pad(c)
        return {"query_tokens": qt.astype(np.int64), "query_pad_mask": qm, "context_tokens": ct.astype(np.int64),
                "context_pad_mask": cm, "block_data": np.asarray([i, i + 1, i, i], dtype=np.int64)}


def model_provider(pre_process=True, post_process=True):
    args = get_args()
    print_rank_0("building BiEncoderModel ...")
    if mpu.get_tensor_model_parallel_world_size() != 1 or mpu.get_pipeline_model_parallel_world_size() != 1:
        raise RuntimeError("model parallel size > 1 is not supported for ICT (as in the reference)")
    config = core_transformer_config_from_args(args)
    return BiEncoderModel(config, args.padded_vocab_size, args.max_position_embeddings, num_tokentypes=2,
                          biencoder_projection_dim=args.biencoder_projection_dim,
                          biencoder_shared_query_context_model=args.biencoder_shared_query_context_model,
                          pre_process=pre_process, post_process=post_process)


def loss_func(output_tensor):
    args = get_args()
    q, c = output_tensor
    loss, stats = ict_loss(q, c, args.hidden_size, args.retriever_score_scaling,
                           args.retriever_report_topk_accuracies or (1,), mpu.get_data_parallel_group())
    # every rank computed the loss of the WHOLE batch; the DP gradient average would divide it by dp again
    return loss * mpu.get_data_parallel_world_size(), {"loss": loss.detach(), **{k: v.detach() for k, v in stats.items()}}


def forward_step(data_iterator, model):
    dev = next(model.parameters()).device
    d = next(data_iterator)
    g = lambda k: d[k].long().to(dev)  # noqa: E731
    qt, ct = g("query_tokens"), g("context_tokens")
    out = model(qt, g("query_pad_mask"), torch.zeros_like(qt), ct, g("context_pad_mask"), torch.zeros_like(ct))
    return out, partial(loss_func)


def train_valid_test_datasets_provider(num_samples):
    args = get_args()
    if getattr(args, "data_path", None) and not args.mock_data:
        from megatronapp_b200.data.ict_dataset import ICTDataset
        from megatronapp_b200.data.indexed_dataset import IndexedDataset

        tok = get_tokenizer()
        blocks = IndexedDataset(args.data_path[0] if len(args.data_path) == 1 else args.data_path[1])
        titles = IndexedDataset(args.titles_data_path) if args.titles_data_path else None
        mk = lambda n, s: ICTDataset(blocks, titles, max(int(n), 1), args.seq_length, args.query_in_block_prob,  # noqa: E731
                                     args.seed + s, tok.cls, tok.sep, tok.pad, use_one_sent_docs=args.use_one_sent_docs)
        return mk(num_samples[0], 0), mk(num_samples[1], 1), mk(num_samples[2], 2)
    mk = lambda n, s: MockICTDataset(max(int(n), 1), args.seq_length, args.padded_vocab_size, s)  # noqa: E731
    return mk(num_samples[0], 1), mk(num_samples[1], 10 ** 6), mk(num_samples[2], 2 * 10 ** 6)


if __name__ == "__main__":
    pretrain(train_valid_test_datasets_provider, model_provider, "encoder_or_decoder", forward_step,
             args_defaults={"tokenizer_type": "NullTokenizer"})
