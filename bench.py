"""Flagship benchmark: GPT-3 6.7B training step, tokens/s for the whole job.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--model gpt3_6.7b]

N > 1 is launched by the driver through ``torch.distributed.run`` (one rank per GPU, NCCL).
Layout per N (BASELINE.json config #2 is the N=8 row): 1 -> TP1, 2 -> TP2+SP, 4 -> TP4+SP,
8 -> TP4 x PP2 + SP.  Weak scaling: ``global_batch = 8 * N`` sequences of 2048 tokens,
micro-batch = TP degree (same per-GPU GEMM shapes at every N).  Synthetic tokens, random-init weights, bf16 compute, fp32 main grads + Adam.

Timed region = exactly K optimizer steps (forward + backward + grad finalize + clip + Adam),
bracketed by barrier + ``cuda.synchronize``, measured with CUDA events, MAX over ranks.
``e2e`` repeats the measurement through the public API with, per step, the H2D copy of that
step's microbatches from pinned host memory and a D2H read of the loss.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MODELS = {
    # name: (layers, hidden, heads, ffn, seq, vocab)
    "gpt3_6.7b": (32, 4096, 32, 16384, 2048, 50257),
    "gpt3_1.3b": (24, 2048, 16, 8192, 2048, 50257),
    "gpt2_125m": (12, 768, 12, 3072, 1024, 50257),
    "tiny": (2, 512, 8, 2048, 256, 4000),
}


def layout_for(n_gpus: int):
    return {1: (1, 1), 2: (2, 1), 4: (4, 1), 8: (4, 2)}.get(n_gpus, (min(n_gpus, 4), max(1, n_gpus // 4)))


def pad_vocab(v: int, tp: int) -> int:
    m = 128 * tp
    return (v + m - 1) // m * m


class ClockSampler:
    """nvidia-smi sampling DURING the timed region (B200_PROFILING.md clocks line)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.proc, self.path = gpu_index, None, f"/tmp/b200_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            p = [x.strip() for x in line.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2]))
            except ValueError:
                continue
            for nm, val in zip(names, p[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if sm:
            out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                   "samples": len(sm)}
        try:
            os.remove(self.path)
        except OSError:
            pass
        return out


def run_own(args):
    import torch
    import torch.distributed as dist

    from megatronapp_b200 import ops
    from megatronapp_b200.bench_model import build_gpt_for_bench, synthetic_batch, train_step
    from megatronapp_b200.parallel import state as ps

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    torch.cuda.set_device(local_rank)
    if world > 1:
        ps.init_distributed("nccl")
    tp, pp = layout_for(args.gpus)
    if args.tp:
        tp, pp = args.tp, args.gpus // args.tp
    L, H, A, F, S, V = MODELS[args.model]
    if args.layers:
        L = args.layers
    vocab = pad_vocab(V, tp)
    # micro-batch = TP degree keeps the per-GPU GEMM shapes of the 1-GPU run (a TP4 rank with
    # mbs 1 would run 35 us GEMMs and be launch-bound); the reference arm uses the same rule
    mbs = args.micro_batch or tp
    dp = args.gpus // (tp * pp)
    global_batch = args.microbatches_per_gpu * args.gpus   # weak scaling: 8 sequences per GPU
    n_mb = global_batch // (mbs * dp)                       # microbatches per data-parallel replica
    ops.ext()
    state = build_gpt_for_bench(L, H, A, S, vocab, mbs, ffn=F, tp=tp, pp=pp, sequence_parallel=tp > 1,
                                device=f"cuda:{local_rank}", lr=1e-4, fused_tp_comm=not args.no_fused_tp_comm,
                                use_distributed_optimizer=dp > 1, dropout=0.0)
    dev_batches = [synthetic_batch(state, f"cuda:{local_rank}", seed=i) for i in range(n_mb)]
    pin_batches = [synthetic_batch(state, f"cuda:{local_rank}", seed=i, pinned=True) for i in range(n_mb)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        train_step(state, dev_batches, sync_loss=False)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ops.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        train_step(state, dev_batches, sync_loss=False)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ops.launch_count()
    clocks = sampler.stop() if rank == 0 else {}
    # end-to-end: pinned-host inputs copied every step + loss read back every step
    last_loss = float("nan")
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.steps):
        last_loss = train_step(state, pin_batches, sync_loss=True)
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    scope_ms = None
    if args.trace_scopes:
        # one extra (untimed) step under the MegaScan tracer: per-scope device time on this rank
        from megatronapp_b200.scan.tracer import get_tracer

        tr = get_tracer()
        tr.configure(True, trace_dir="/tmp/b200_bench_trace", interval=1, continuous_iterations=1)
        tr.log = lambda: None
        tr.iteration_begin(0)
        train_step(state, dev_batches, sync_loss=False)
        tr.iteration_end()
        tot, cnt, stack = {}, {}, []
        for r in tr._records:
            if r["ph"] == "B":
                stack.append(r)
            elif r["ph"] == "E":
                for k in range(len(stack) - 1, -1, -1):
                    if stack[k]["name"] == r["name"]:
                        b = stack.pop(k)
                        tot[r["name"]] = tot.get(r["name"], 0.0) + (r["rel_ts"] - b["rel_ts"]) / 1e6
                        cnt[r["name"]] = cnt.get(r["name"], 0) + 1
                        break
        scope_ms = {k: {"ms": round(v, 3), "calls": cnt[k]} for k, v in sorted(tot.items(), key=lambda kv: -kv[1])}
        tr.configure(False)
        barrier()
    last_loss = _finite_from_any_rank(last_loss, world)
    t = torch.tensor([ms, ms_e2e], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    tokens = global_batch * S * args.steps
    first_or_last = state.pre_process or state.post_process
    h2d = sum(v.numel() * v.element_size() for b in pin_batches for v in b.values()) if first_or_last else 0
    if rank == 0:
        par = f"tp{tp}" + (f"pp{pp}" if pp > 1 else "") + (f"dp{dp}" if dp > 1 else "") + ("+sp" if tp > 1 else "")
        out = {
            "metric": "GPT training tokens/sec (whole job, device-timed, max over ranks)",
            "value": tokens / (ms / 1e3), "unit": "tokens/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic tokens, random-init weights",
            "impl": "megatronapp_b200",
            "config": {"model": args.model, "layers": L, "hidden": H, "heads": A, "ffn": F, "vocab_padded": vocab,
                       "global_batch": global_batch, "micro_batch": mbs, "seq_len": S, "parallelism": par,
                       "optimizer": "Adam fp32 master + fp32 main grads, clip 1.0",
                       "l2_policy": "working set (>=13 GB weights + grads per step) >> 126 MB L2; no explicit flush",
                       "attention_impl": "native" if ops.native_attention_available() else "sdpa-library",
                       "fused_tp_comm": (tp > 1 and not args.no_fused_tp_comm)},
            "clocks": clocks,
            "e2e": {"value": tokens / (ms_e2e / 1e3), "unit": "tokens/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4 if state.post_process else 0,
                    "final_loss": last_loss},
            "gpu_launches": launches,
            "library_fallbacks": dict(ops.fallback_counts),
            "tflops_per_gpu": 6.0 * _param_count(L, H, F, vocab) * tokens / (ms / 1e3) / 1e12 / args.gpus,
        }
        if scope_ms is not None:
            out["rank0_scope_ms_one_step"] = scope_ms
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _finite_from_any_rank(x, world):
    """The loss lives on the last pipeline stage; rank 0 prints, so fetch a finite copy."""
    import math

    import torch.distributed as dist

    if world <= 1:
        return x
    vals = [None] * world
    dist.all_gather_object(vals, x)
    for v in reversed(vals):
        if v is not None and isinstance(v, float) and math.isfinite(v):
            return v
    return x


def _param_count(L, H, F, V):
    return L * (4 * H * H + 2 * H * F) + V * H


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--model", default="gpt3_6.7b", choices=list(MODELS))
    ap.add_argument("--layers", type=int, default=None, help="debug only: override layer count (invalidates the number)")
    ap.add_argument("--micro-batch", type=int, default=None, help="default: TP degree")
    ap.add_argument("--microbatches-per-gpu", type=int, default=8)
    ap.add_argument("--tp", type=int, default=None)
    ap.add_argument("--no-fused-tp-comm", action="store_true", help="NCCL collectives instead of fused AG/RS GEMMs")
    ap.add_argument("--trace-scopes", action="store_true", help="add one untimed traced step; per-scope device ms in the JSON")
    ap.add_argument("--fused-rs", action="store_true", help="also fuse GEMM->reduce-scatter (default: AG->GEMM only)")
    args = ap.parse_args()
    if args.fused_rs:
        os.environ["B200_FUSED_RS"] = "1"
    if args.impl == "reference":
        from baseline.reference_arm import run_reference

        run_reference(args, MODELS, layout_for, pad_vocab)
    else:
        run_own(args)


if __name__ == "__main__":
    main()
