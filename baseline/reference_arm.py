"""``bench.py --impl reference``: the UNMODIFIED reference (megatron.core 0.12.2 installed from
/root/reference into baseline/_ref with pip --no-deps) on the same metric/config.  See DESIGN.md
for the install record."""
from __future__ import annotations

import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def unavailable(why: str):
    print(json.dumps({"impl": "reference", "unavailable": why}), flush=True)
    sys.exit(0)


def run_reference(args, MODELS, layout_for, pad_vocab):
    if not os.path.isdir(os.path.join(REF, "megatron", "core")):
        unavailable("baseline/_ref/megatron/core not present (reference not installed)")
    if args.model in ("mixtral_8x7b", "tiny_moe"):
        unavailable("MoE with tensor parallelism requires sequence parallelism, which the reference's stock local spec "
                    "(torch norm, no Apex/TE offline) refuses; the in-tree comparator is bench.py --moe-baseline "
                    "(NCCL all-to-all dispatcher + per-expert modules, the reference's algorithm)")
    if args.fbd or args.dpp or args.fp8:
        unavailable("the reference arm drives the stock schedule only: MegaDPP needs the reference's ibverbs "
                    "shm_tensor_new_rdma extensions and fp8 needs TransformerEngine (neither builds offline), MegaFBD needs "
                    "its thread-based virtual-TP shim -- comparator: the same config without the flag")
    try:
        from .reference_runner import main as ref_main
    except Exception as e:  # noqa: BLE001
        unavailable(f"reference runner import failed: {type(e).__name__}: {str(e)[:160]}")
    ref_main(args, MODELS, layout_for, pad_vocab)
