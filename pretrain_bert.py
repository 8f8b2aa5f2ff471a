"""Pretrain BERT (entry point kept from the reference ``pretrain_bert.py``: MLM + optional NSP)."""
import os

if not os.path.exists("/dev/nvidia0"):   # CPU/gloo mode: avoid OpenMP oversubscription
    os.environ.setdefault("OMP_NUM_THREADS", "2")

from functools import partial

import torch
import torch.nn.functional as F

from megatronapp_b200.models.bert import BertModel, bert_layer_local_spec
from megatronapp_b200.parallel import state as mpu
from megatronapp_b200.training.arguments import core_transformer_config_from_args
from megatronapp_b200.training.global_vars import get_args
from megatronapp_b200.training.training import pretrain
from megatronapp_b200.training.utils import broadcast_batch, print_rank_0


class MockBertDataset(torch.utils.data.Dataset):
    def __init__(self, n, seq, vocab, seed, mask_prob=0.15, binary_head=True):
        self.n, self.seq, self.vocab, self.seed, self.mask_prob, self.binary = n, seq, vocab, seed, mask_prob, binary_head

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        g = torch.Generator().manual_seed(self.seed + i)
        text = torch.randint(5, self.vocab, (self.seq,), generator=g)
        mask_pos = torch.rand(self.seq, generator=g) < self.mask_prob
        labels = torch.where(mask_pos, text, torch.full_like(text, -1))
        inp = torch.where(mask_pos, torch.full_like(text, 4), text)   # 4 = [MASK]
        return {"text": inp, "types": (torch.arange(self.seq) >= self.seq // 2).long(), "labels": labels.clamp_min(0),
                "is_random": torch.randint(0, 2, (1,), generator=g)[0], "loss_mask": mask_pos.float(),
                "padding_mask": torch.ones(self.seq, dtype=torch.long)}


def model_provider(pre_process=True, post_process=True):
    args = get_args()
    print_rank_0("building BERT model ...")
    config = core_transformer_config_from_args(args)
    if args.use_legacy_models:
        from megatronapp_b200.legacy.model import BertModel as LegacyBertModel

        return LegacyBertModel(config, num_tokentypes=2 if args.bert_binary_head else 0,
                               add_binary_head=args.bert_binary_head, parallel_output=True, pre_process=pre_process,
                               post_process=post_process)
    return BertModel(config=config, num_tokentypes=2 if args.bert_binary_head else 0,
                     transformer_layer_spec=bert_layer_local_spec(), vocab_size=args.padded_vocab_size,
                     max_sequence_length=args.max_position_embeddings, pre_process=pre_process,
                     post_process=post_process, add_binary_head=args.bert_binary_head, parallel_output=True,
                     share_embeddings_and_output_weights=not args.untie_embeddings_and_output_weights)


def loss_func(loss_mask, sentence_order, output_tensor):
    lm_loss_, sop_logits = output_tensor
    lm_loss_ = lm_loss_.float()
    loss_mask = loss_mask.float()
    lm_loss = torch.sum(lm_loss_.view(-1) * loss_mask.reshape(-1)) / loss_mask.sum().clamp_min(1)
    if sop_logits is not None:
        sop_loss = F.cross_entropy(sop_logits.view(-1, 2).float(), sentence_order.view(-1), ignore_index=-1)
        return lm_loss + sop_loss, {"lm loss": lm_loss.detach(), "sop loss": sop_loss.detach()}
    return lm_loss, {"lm loss": lm_loss.detach()}


def forward_step(data_iterator, model):
    dev = next(model.parameters()).device
    d = broadcast_batch(data_iterator, ("text", "types", "labels", "is_random", "padding_mask"), ("loss_mask",))
    tokens, types, labels = d["text"].to(dev), d["types"].to(dev), d["labels"].to(dev)
    out = model(tokens, d["padding_mask"].to(dev), tokentype_ids=types, lm_labels=labels)
    return out, partial(loss_func, d["loss_mask"].to(dev), d["is_random"].to(dev))


def train_valid_test_datasets_provider(num_samples):
    args = get_args()
    if getattr(args, "data_path", None) and not args.mock_data:
        # sentence-level .bin/.idx corpus (tools/preprocess_data.py --split-sentences) -> masked LM samples
        from megatronapp_b200.data.gpt_dataset import parse_split, split_document_ranges
        from megatronapp_b200.data.indexed_dataset import IndexedDataset
        from megatronapp_b200.data.masked_dataset import BERTMaskedWordPieceDataset, MaskedWordPieceDatasetConfig
        from megatronapp_b200.training.global_vars import get_tokenizer

        prefix = args.data_path[0] if len(args.data_path) == 1 else args.data_path[1]
        idx = IndexedDataset(prefix)
        cfg = MaskedWordPieceDatasetConfig(sequence_length=args.seq_length, masking_probability=args.mask_prob,
                                           short_sequence_probability=args.short_seq_prob, random_seed=args.seed,
                                           classification_head=args.bert_binary_head)
        ranges = split_document_ranges(len(idx.document_indices) - 1, parse_split(args.split))
        tok = get_tokenizer()
        return tuple(BERTMaskedWordPieceDataset(idx, max(int(n), 1), cfg, tok, doc_range=r) if r[1] > r[0] else None
                     for n, r in zip(num_samples, ranges))
    mk = lambda n, s: MockBertDataset(max(int(n), 1), args.seq_length, args.padded_vocab_size, s, args.mask_prob, args.bert_binary_head)  # noqa: E731
    return mk(num_samples[0], 1), mk(num_samples[1], 2), mk(num_samples[2], 3)


if __name__ == "__main__":
    pretrain(train_valid_test_datasets_provider, model_provider, "encoder_or_decoder", forward_step,
             args_defaults={"tokenizer_type": "NullTokenizer"})
