"""Pretrain T5 (entry point kept from the reference ``pretrain_t5.py``)."""
import os

if not os.path.exists("/dev/nvidia0"):   # CPU/gloo mode: avoid OpenMP oversubscription
    os.environ.setdefault("OMP_NUM_THREADS", "2")

from functools import partial

import torch

from megatronapp_b200.models.T5 import T5Model, get_t5_decoder_block_spec, get_t5_encoder_block_spec
from megatronapp_b200.parallel import state as mpu
from megatronapp_b200.training.arguments import core_transformer_config_from_args
from megatronapp_b200.training.global_vars import get_args
from megatronapp_b200.training.training import pretrain
from megatronapp_b200.training.utils import broadcast_batch, print_rank_0


class MockT5Dataset(torch.utils.data.Dataset):
    def __init__(self, n, enc_seq, dec_seq, vocab, seed):
        self.n, self.es, self.ds, self.vocab, self.seed = n, enc_seq, dec_seq, vocab, seed

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        g = torch.Generator().manual_seed(self.seed + i)
        enc = torch.randint(3, self.vocab, (self.es,), generator=g)
        dec = torch.randint(3, self.vocab, (self.ds + 1,), generator=g)
        return {"text_enc": enc, "text_dec": dec[:-1], "labels": dec[1:], "loss_mask": torch.ones(self.ds),
                "enc_mask": torch.ones(self.es, self.es), "dec_mask": torch.tril(torch.ones(self.ds, self.ds)),
                "enc_dec_mask": torch.ones(self.ds, self.es)}


def model_provider(pre_process=True, post_process=True, add_encoder=True, add_decoder=True):
    args = get_args()
    print_rank_0("building T5 model ...")
    config = core_transformer_config_from_args(args)
    n_enc = args.encoder_num_layers or args.num_layers
    n_dec = args.decoder_num_layers or args.num_layers
    pp, split = mpu.get_pipeline_model_parallel_world_size(), mpu.get_pipeline_model_parallel_split_rank()
    if pp > 1:                       # encoder layers over stages [0, split), decoder layers over [split, pp)
        if n_enc % split or n_dec % (pp - split):
            raise ValueError(f"encoder/decoder layers ({n_enc}/{n_dec}) must divide by their stage counts ({split}/{pp - split})")
        n_enc, n_dec = n_enc // split, n_dec // (pp - split)
    return T5Model(config, config, get_t5_encoder_block_spec(n_enc), get_t5_decoder_block_spec(n_dec),
                   vocab_size=args.padded_vocab_size, max_sequence_length=args.max_position_embeddings,
                   pre_process=pre_process, post_process=post_process, parallel_output=True,
                   share_embeddings_and_output_weights=not args.untie_embeddings_and_output_weights,
                   position_embedding_type=args.position_embedding_type,
                   relative_attention_num_buckets=args.relative_attention_num_buckets,
                   relative_attention_max_distance=args.relative_attention_max_distance,
                   add_encoder=add_encoder, add_decoder=add_decoder)


def loss_func(loss_mask, output_tensor):
    l = output_tensor.float()
    loss = torch.sum(l.view(-1) * loss_mask.reshape(-1)) / loss_mask.sum()
    return loss, {"lm loss": loss.detach()}


def forward_step(data_iterator, model):
    dev = next(model.parameters()).device
    d = broadcast_batch(data_iterator, ("text_enc", "text_dec", "labels", "enc_mask", "dec_mask", "enc_dec_mask"), ("loss_mask",))
    d = {k: v.to(dev) for k, v in d.items()}
    out = model(d["text_enc"], d["text_dec"], d["enc_mask"], d["dec_mask"], d["enc_dec_mask"], lm_labels=d["labels"])
    return out, partial(loss_func, d["loss_mask"])


def train_valid_test_datasets_provider(num_samples):
    args = get_args()
    dec = args.decoder_seq_length or max(8, args.seq_length // 4)
    if getattr(args, "data_path", None) and not args.mock_data:
        # sentence-level .bin/.idx corpus -> span-corruption samples; the sentinels are the tokenizer's extra ids
        # (--vocab-extra-ids 100), as in the reference's T5MaskedWordPieceDataset
        from megatronapp_b200.data.gpt_dataset import parse_split, split_document_ranges
        from megatronapp_b200.data.indexed_dataset import IndexedDataset
        from megatronapp_b200.data.masked_dataset import MaskedWordPieceDatasetConfig, T5MaskedWordPieceDataset
        from megatronapp_b200.training.global_vars import get_tokenizer

        prefix = args.data_path[0] if len(args.data_path) == 1 else args.data_path[1]
        idx = IndexedDataset(prefix)
        tok = get_tokenizer()
        cfg = MaskedWordPieceDatasetConfig(sequence_length=args.encoder_seq_length or args.seq_length,
                                           sequence_length_decoder=dec, masking_probability=args.mask_prob,
                                           short_sequence_probability=args.short_seq_prob, random_seed=args.seed,
                                           classification_head=False)
        n_extra = getattr(args, "vocab_extra_ids", 0) or 0
        sentinels = list(range(tok.vocab_size - n_extra, tok.vocab_size)) if n_extra else None
        ranges = split_document_ranges(len(idx.document_indices) - 1, parse_split(args.split or "969,30,1"))
        return tuple(T5MaskedWordPieceDataset(idx, max(int(n), 1), cfg, tok, doc_range=r, sentinel_tokens=sentinels)
                     if r[1] > r[0] else None for n, r in zip(num_samples, ranges))
    mk = lambda n, s: MockT5Dataset(max(int(n), 1), args.seq_length, dec, args.padded_vocab_size, s)  # noqa: E731
    return mk(num_samples[0], args.seed), mk(num_samples[1], args.seed + 1), mk(num_samples[2], args.seed + 2)


if __name__ == "__main__":
    pretrain(train_valid_test_datasets_provider, model_provider, "encoder_and_decoder", forward_step,
             args_defaults={"tokenizer_type": "NullTokenizer"})
