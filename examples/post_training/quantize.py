#!/usr/bin/env python
"""PTQ of a GPT checkpoint (reference ``examples/post_training/modelopt/quantize.py``).

    torchrun --nproc-per-node 1 examples/post_training/quantize.py <model args> --load ck --save ck_fp8 \\
        --export-quant-cfg fp8 --calib-size 64 [--export-trtllm-dir out]

Calibrates on ``--data-path`` when given (first ``--calib-size`` training samples), else on synthetic tokens; writes
the folded weights as a normal checkpoint plus ``quantizer_state.pt`` (amax / pre-quant scales) and, optionally, the
real low-precision payloads (``quantized_weights.pt``).
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

from megatronapp_b200.post_training import quantization as ptq  # noqa: E402
from megatronapp_b200.training import checkpointing  # noqa: E402
from megatronapp_b200.training.global_vars import get_args  # noqa: E402
from megatronapp_b200.training.initialize import initialize_megatron  # noqa: E402
from megatronapp_b200.training.training import get_model  # noqa: E402
from megatronapp_b200.training.utils import print_rank_0  # noqa: E402
from pretrain_gpt import model_provider  # noqa: E402


def _extra_args(parser):
    g = parser.add_argument_group("ptq")
    g.add_argument("--export-quant-cfg", default="fp8", choices=list(ptq.QUANT_CFG_CHOICES))
    g.add_argument("--calib-size", type=int, default=64)
    g.add_argument("--calib-batch-size", type=int, default=4)
    g.add_argument("--export-real-weights", action="store_true")
    return parser


def main():
    initialize_megatron(extra_args_provider=_extra_args,
                        args_defaults={"tokenizer_type": "NullTokenizer", "micro_batch_size": 1, "no_load_optim": True})
    args = get_args()
    model = get_model(model_provider, wrap_with_ddp=False)
    if args.load:
        checkpointing.load_checkpoint(model, None, None)
    m = model[0]
    dev = next(m.parameters()).device
    g = torch.Generator().manual_seed(args.seed)
    batches = [torch.randint(1, args.padded_vocab_size - 1, (args.calib_batch_size, args.seq_length), generator=g)
               for _ in range(max(1, args.calib_size // args.calib_batch_size))]
    pos = torch.arange(args.seq_length, device=dev).unsqueeze(0).expand(args.calib_batch_size, -1)

    def loop(mod):
        with torch.no_grad():
            for b in batches:
                mod(b.to(dev), pos, None)

    ptq.quantize(m, args.export_quant_cfg, loop)
    state = ptq.quantizer_state_dict(m)
    print_rank_0(f"quantised with {args.export_quant_cfg}: {len(state)} quantiser tensors")
    if args.save:
        checkpointing.save_checkpoint(0, model, None, None, 0)
        if torch.distributed.get_rank() == 0:
            torch.save({"cfg": args.export_quant_cfg, "state": state}, os.path.join(args.save, "quantizer_state.pt"))
            if args.export_real_weights:
                torch.save(ptq.export_quantized_weights(m), os.path.join(args.save, "quantized_weights.pt"))


if __name__ == "__main__":
    main()
