"""Supervised fine-tuning of a Retro model on QA / instruction data with retrieved passages (reference
``tools/retro/sft/sft_retro.py``).

    python tools/retro/sft/sft_retro.py --load retro_ckpt --finetune --train-data-path qa_train.json --valid-data-path qa_dev.json \\
        --retro-chunk-length 64 --retro-num-neighbors 2 --retro-retrieved-length 128 --seq-length 2048 ... (model / optimizer flags)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))

import torch  # noqa: E402

import pretrain_retro as base  # noqa: E402
from megatronapp_b200.training.global_vars import get_args, get_tokenizer  # noqa: E402
from megatronapp_b200.training.training import pretrain  # noqa: E402
from tools.retro.sft.dataset_conv import RetroSFTDataset  # noqa: E402


def add_sft_args(parser):
    parser = base.add_retro_args(parser)
    g = parser.add_argument_group("retro sft")
    g.add_argument("--train-data-path", nargs="+", default=None, help="QA json / jsonl with question, answers, ctxs")
    g.add_argument("--valid-data-path", nargs="+", default=None)
    g.add_argument("--ft-neighbours", type=int, default=0, help="also prepend this many passages to the prompt as plain text")
    g.add_argument("--loss-on-prompt", action="store_true", help="train on the prompt tokens too (default: answer only)")
    return parser


class _Epochs(torch.utils.data.Dataset):
    """the finite SFT set repeated to the number of samples the run will consume (shuffled differently every epoch)"""

    def __init__(self, ds, n, seed=1234):
        self.ds, self.n, self.seed = ds, max(int(n), len(ds)), seed

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        epoch, j = divmod(i, len(self.ds))
        perm = torch.randperm(len(self.ds), generator=torch.Generator().manual_seed(self.seed + epoch))
        return self.ds[int(perm[j])]


def train_valid_test_datasets_provider(num_samples):
    a, tok = get_args(), get_tokenizer()
    if not a.train_data_path:
        raise ValueError("--train-data-path is required")
    mk = lambda p: RetroSFTDataset(p, tok, a.seq_length, a.retro_chunk_length, a.retro_num_neighbors, a.retro_retrieved_length,  # noqa: E731
                                   a.ft_neighbours, answer_loss_only=not a.loss_on_prompt)
    train = mk(a.train_data_path[0])
    valid = mk(a.valid_data_path[0]) if a.valid_data_path else train
    return _Epochs(train, num_samples[0], a.seed), _Epochs(valid, num_samples[1], a.seed + 1), _Epochs(valid, num_samples[2], a.seed + 2)


if __name__ == "__main__":
    train_valid_test_datasets_provider.is_distributed = True
    pretrain(train_valid_test_datasets_provider, base.model_provider, "encoder_or_decoder", base.forward_step,
             extra_args_provider=add_sft_args, args_defaults={"tokenizer_type": "GPT2BPETokenizer"})
