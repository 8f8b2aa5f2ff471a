#!/usr/bin/env python
"""Change the TP / PP layout of a Mamba or hybrid (Mamba + attention + MLP) ``mp_rank_*`` checkpoint.

Parity: reference ``tools/checkpoint/hybrid_conversion.py``.  The Mamba mixer's ``in_proj`` and depth-wise conv are
sharded *per logical segment* (``[z | x | B | C | dt]`` resp. ``[x | B | C]`` of each rank's heads / groups), so a
plain chunk along dim 0 would mix segments: every segment is merged / split on its own.  Attention and MLP layers
reuse the rules of ``convert.py``.

    python tools/checkpoint/hybrid_conversion.py --load-dir ck_tp1 --save-dir ck_tp2pp2 \\
        --target-tp-size 2 --target-pp-size 2
"""
from __future__ import annotations

import argparse
import os
import sys
from types import SimpleNamespace
from typing import Dict, List

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import convert as conv  # noqa: E402

MAMBA_DIM0 = ("mixer.dt_bias", "mixer.A_log", "mixer.D", "mixer.norm_weight")


def mamba_dims(args) -> SimpleNamespace:
    d_model = args.hidden_size
    d_inner = 2 * d_model
    headdim = getattr(args, "mamba_head_dim", 64)
    nheads = d_inner // headdim
    ngroups = min(getattr(args, "mamba_num_groups", 8), nheads)
    return SimpleNamespace(d_inner=d_inner, nheads=nheads, ngroups=ngroups, d_state=getattr(args, "mamba_state_dim", 128))


def _segments(name: str, d: SimpleNamespace) -> List[int]:
    """Full (unsharded) row counts of the segments of a segmented tensor; ``[]`` when ``name`` is not segmented."""
    gn = d.ngroups * d.d_state
    if name.endswith("mixer.in_proj.weight"):
        return [d.d_inner, d.d_inner, gn, gn, d.nheads]
    if name.endswith(("mixer.conv1d.weight", "mixer.conv1d.bias")):
        return [d.d_inner, gn, gn]
    return []


def merge(name: str, shards: List[torch.Tensor], d: SimpleNamespace, glu: bool) -> torch.Tensor:
    tp = len(shards)
    segs = _segments(name, d)
    if tp == 1:
        return shards[0]
    if segs:
        parts = [torch.split(s, [n // tp for n in segs], dim=0) for s in shards]
        return torch.cat([torch.cat([p[i] for p in parts], dim=0) for i in range(len(segs))], dim=0)
    if name.endswith(MAMBA_DIM0):
        return torch.cat(shards, dim=0)
    if name.endswith("mixer.out_proj.weight"):
        return torch.cat(shards, dim=1)
    return conv.merge_tp(name, shards, glu)


def split(name: str, full: torch.Tensor, tp: int, d: SimpleNamespace, glu: bool) -> List[torch.Tensor]:
    segs = _segments(name, d)
    if tp == 1:
        return [full.clone()]
    if segs:
        if any(n % tp for n in segs):
            raise ValueError(f"{name}: segments {segs} do not divide by tp={tp}")
        parts = [torch.chunk(p, tp, dim=0) for p in torch.split(full, segs, dim=0)]
        return [torch.cat([p[r] for p in parts], dim=0).contiguous() for r in range(tp)]
    if name.endswith(MAMBA_DIM0):
        return [c.contiguous() for c in torch.chunk(full, tp, dim=0)]
    if name.endswith("mixer.out_proj.weight"):
        return [c.contiguous() for c in torch.chunk(full, tp, dim=1)]
    return conv.split_tp(name, full, tp, glu)


def load(load_dir: str) -> Dict:
    from megatronapp_b200.training import checkpointing as ck

    with open(ck.get_checkpoint_tracker_filename(load_dir)) as f:
        txt = f.read().strip()
    it = 0 if txt == "release" else int(txt)
    base = ck.get_checkpoint_name(load_dir, it, release=(it == 0), return_base_dir=True)
    dirs = sorted(x for x in os.listdir(base) if x.startswith("mp_rank_"))
    tp = 1 + max(int(x.split("_")[2]) for x in dirs)
    pp = 1 + max((int(x.split("_")[3]) for x in dirs if len(x.split("_")) > 3), default=0)
    shards, args = {}, None
    for x in dirs:
        parts = x.split("_")
        t, p = int(parts[2]), (int(parts[3]) if len(parts) > 3 else 0)
        sd = torch.load(os.path.join(base, x, "model_optim_rng.pt"), map_location="cpu", weights_only=False)
        shards[(t, p)] = sd["model"]
        args = args or sd.get("args")
    d = mamba_dims(args)
    glu = bool(getattr(args, "swiglu", False))
    lps = conv.layers_per_stage(args.num_layers, pp)
    weights: Dict[str, torch.Tensor] = {}
    for p in range(pp):
        for name in shards[(0, p)]:
            if name.endswith("_extra_state"):
                continue
            full = merge(name, [shards[(t, p)][name] for t in range(tp)], d, glu)
            m = conv._LAYER.match(name)
            g = f"{m.group(1)}{int(m.group(2)) + p * lps}{m.group(3)}" if m else name
            weights.setdefault(g, full)
    return {"weights": weights, "args": args, "iteration": it, "dims": d, "glu": glu}


def save(ckpt: Dict, save_dir: str, tp: int, pp: int):
    from megatronapp_b200.training import checkpointing as ck

    args, d, glu, it = ckpt["args"], ckpt["dims"], ckpt["glu"], ckpt["iteration"]
    lps = conv.layers_per_stage(args.num_layers, pp)
    new_args = SimpleNamespace(**{**vars(args), "tensor_model_parallel_size": tp, "pipeline_model_parallel_size": pp})
    for p in range(pp):
        stage = {}
        for name, w in ckpt["weights"].items():
            m = conv._LAYER.match(name)
            if m:
                li = int(m.group(2))
                if li // lps != p:
                    continue
                lname = f"{m.group(1)}{li - p * lps}{m.group(3)}"
            elif name.startswith("embedding."):
                if p != 0:
                    continue
                lname = name
            else:
                if p != pp - 1:
                    continue
                lname = name
            stage[lname] = split(name, w, tp, d, glu)
        for t in range(tp):
            path = ck.get_checkpoint_name(save_dir, it, release=(it == 0), pipeline_parallel=pp > 1, tensor_rank=t,
                                          pipeline_rank=p, expert_parallel=False, expert_rank=0)
            os.makedirs(os.path.dirname(path), exist_ok=True)
            torch.save({"model": {k: (v[t] if len(v) > 1 else v[0]) for k, v in stage.items()}, "args": new_args,
                        "iteration": it, "checkpoint_version": 3.0}, path)
    with open(ck.get_checkpoint_tracker_filename(save_dir), "w") as f:
        f.write(str(it) if it else "release")


def main(argv=None):
    ap = argparse.ArgumentParser(description="reshard a Mamba / hybrid checkpoint")
    ap.add_argument("--load-dir", required=True)
    ap.add_argument("--save-dir", required=True)
    ap.add_argument("--target-tp-size", type=int, default=1)
    ap.add_argument("--target-pp-size", type=int, default=1)
    a = ap.parse_args(argv)
    ck = load(a.load_dir)
    save(ck, a.save_dir, a.target_tp_size, a.target_pp_size)
    print(f"resharded {len(ck['weights'])} tensors -> tp={a.target_tp_size} pp={a.target_pp_size} at {a.save_dir}")


if __name__ == "__main__":
    main()
