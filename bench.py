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
    "mixtral_8x7b": (32, 4096, 32, 14336, 4096, 32000),
    "llama3_8b": (32, 4096, 32, 14336, 8192, 128256),
    "tiny": (2, 512, 8, 2048, 256, 4000),
    "tiny_moe": (2, 512, 8, 1024, 256, 4000),
}
# architecture switches beyond the GPT-3 defaults (gelu MLP, LayerNorm, learned positions, biases)
ARCH = {
    "mixtral_8x7b": dict(swiglu=True, rope=True, rmsnorm=True, kv_groups=8, bias=False, num_experts=8, moe_topk=2),
    "llama3_8b": dict(swiglu=True, rope=True, rmsnorm=True, kv_groups=8, bias=False),
    "tiny_moe": dict(swiglu=True, rope=True, rmsnorm=True, kv_groups=2, bias=False, num_experts=4, moe_topk=2),
}


def layout_for(n_gpus: int):
    return {1: (1, 1), 2: (2, 1), 4: (4, 1), 8: (4, 2)}.get(n_gpus, (min(n_gpus, 4), max(1, n_gpus // 4)))


def parse_layout(spec: str):
    """'tp2ep4', 'tp1pp4dp2', 'tp8' ... -> dict(tp, pp, ep, dp?)"""
    import re

    out = {"tp": 1, "pp": 1, "ep": 1}
    for k, v in re.findall(r"(tp|pp|ep|dp|etp|vpp)(\d+)", spec):
        out[k] = int(v)
    return out


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
    ep, etp, vpp = 1, None, args.vpp
    if args.layout:
        lay = parse_layout(args.layout)
        tp, pp, ep = lay["tp"], lay["pp"], lay["ep"]
        etp, vpp = lay.get("etp"), lay.get("vpp", vpp)
    L, H, A, F, S, V = MODELS[args.model]
    arch = dict(ARCH.get(args.model, {}))
    if args.layers:
        L = args.layers
    if args.seq_len:
        S = args.seq_len
    vocab = pad_vocab(V, tp)
    # micro-batch = TP degree keeps the per-GPU GEMM shapes of the 1-GPU run (a TP4 rank with
    # mbs 1 would run 35 us GEMMs and be launch-bound); the reference arm uses the same rule
    mbs = args.micro_batch or tp
    fb = 2 if args.fbd else 1                               # MegaFBD: every data-parallel slot is a (forward, backward) pair
    dp = args.gpus // (tp * pp * fb)
    global_batch = args.global_batch or args.microbatches_per_gpu * args.gpus   # weak scaling: 8 sequences per GPU
    n_mb = global_batch // (mbs * dp)                       # microbatches per data-parallel replica
    ops.ext()
    moe = arch.get("num_experts") is not None
    if moe:
        arch.update(ep=ep, etp=etp, moe_grouped=not args.moe_baseline, moe_ffn=F,
                    moe_dispatcher="alltoall" if args.moe_baseline else "flex", moe_aux_loss_coeff=1e-2)
    recompute_layers, bf16_moments = args.recompute_layers, args.bf16_adam_moments
    if args.model == "mixtral_8x7b" and args.gpus <= 8 and not args.layers:
        # 46.7 B parameters on <= 8 x 180 GB: fp32 main grads + fp32 master weights stay, but the Adam moments are kept in
        # bf16 (precision-aware optimizer) and every layer is recomputed in backward -- the round-2 runs with fp32 moments
        # went out of memory at 169 GB even with half of the layers recomputed (profiles/README.md)
        recompute_layers = recompute_layers or L // pp
        bf16_moments = True
    state = build_gpt_for_bench(L, H, A, S, vocab, mbs, ffn=F, tp=tp, pp=pp, sequence_parallel=tp > 1,
                                device=f"cuda:{local_rank}", lr=1e-4, fused_tp_comm=not args.no_fused_tp_comm,
                                use_distributed_optimizer=dp > 1, overlap_grad_reduce=(dp > 1 and not args.no_dp_overlap),
                                fused_grad_rs=(args.fused_grad_rs and dp > 1 and not moe), dropout=0.0, vpp=vpp, fbd=args.fbd,
                                fp8="hybrid" if args.fp8 else None, fp8_recipe=args.fp8_recipe,
                                recompute_layers=recompute_layers, bf16_adam_moments=bf16_moments, **arch)
    if args.fbd:
        from megatronapp_b200.fbd.schedule import install_fbd_schedule

        install_fbd_schedule()
    if args.dpp:
        from megatronapp_b200.dpp.scheduler import install_dpp_schedule

        install_dpp_schedule(policy=args.dpp)
    slow = None
    if args.slow_rank_inject:
        r_, cyc = args.slow_rank_inject.split(":")
        slow = (int(r_), int(cyc))
        if rank == slow[0]:
            # MegaScan software slow-rank injector: burns CYCLES SM clocks at two points of every transformer layer, forward
            # and backward, on this rank only -- what a clock-capped GPU looks like to its tensor-parallel peers
            ops.set_slow_rank_spin(slow[1])
    _train_step = train_step

    def train_step(st, batches, sync_loss=True):            # noqa: F811 - add-on aware wrapper around the public API
        if args.fbd:
            from megatronapp_b200.fbd.schedule import sync_forward_weights
            from megatronapp_b200.parallel import state as ps_

            out = _train_step(st, batches, sync_loss=sync_loss, do_optimizer=ps_.is_backward_stage())
            sync_forward_weights(st.model)
            return out
        return _train_step(st, batches, sync_loss=sync_loss)

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
    # ---- MegaScan metrics (BASELINE.json): the same K steps again under the tracer (CUDA-event scopes, per-iteration
    # hand-off to the resolver thread, records written by each rank's background writer) -> trace overhead %;
    # the scope records give the exposed (compute-stream-blocking) communication time per step.
    from megatronapp_b200.scan.tracer import get_tracer

    tr = get_tracer()
    # one directory for the whole job (every rank writes its own file into it): the launcher's pid is common to all ranks
    trace_dir = os.path.join(os.environ.get("B200_BENCH_TRACE_DIR", "/tmp"),
                             f"b200_bench_trace_{os.environ.get('MASTER_PORT', '0')}_{os.getppid() if world > 1 else os.getpid()}")
    tr.configure(True, trace_dir=trace_dir, interval=args.steps, continuous_iterations=args.steps)
    scope_stats = {}
    kept = []
    orig_log = tr.log

    def _log_and_keep():
        tr._drain()
        kept.extend(tr._records)       # the bench's own analysis (below, outside the timed region) reads the same records
        orig_log()

    tr.log = _log_and_keep
    barrier()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for i in range(args.steps):
        tr.iteration_begin(i)
        train_step(state, dev_batches, sync_loss=False)
        tr.iteration_end()
    tr._ship_at = None
    tr.log()          # drain the resolver + hand the window to the writer thread: inside the timed region
    g1.record()
    barrier()
    ms_traced = g0.elapsed_time(g1)
    scope_stats.update(_digest_scopes(kept))
    tr.log = orig_log
    try:
        tr.shutdown()            # flush the writer thread while trace_dir still points at this run's directory
    except Exception:
        pass
    tr.configure(False, trace_dir=trace_dir)
    detector = None
    if args.detect:
        barrier()                  # every rank's file is final
        if rank == 0:
            from megatronapp_b200.scan import aggregate as _agg

            ab = os.path.join(trace_dir, "abnormal.txt")
            _, findings = _agg.aggregate(trace_dir, os.path.join(trace_dir, "aggregated.json"), detect=True, abnormal_file=ab)
            detector = {"abnormal_txt": open(ab).read().strip() if os.path.exists(ab) else None,
                        "injected_rank": slow[0] if slow else None, "findings": findings[:4]}
    if not os.environ.get("B200_BENCH_TRACE_DIR"):
        barrier()
        if rank == 0:
            import shutil

            shutil.rmtree(trace_dir, ignore_errors=True)
    exposed_ms = scope_stats.get("exposed_comm_ns", 0) / 1e6 / max(1, args.steps)
    scope_ms = {k: {"ms": round(v["ns"] / 1e6 / args.steps, 3), "calls": v["calls"] // max(1, args.steps)}
                for k, v in sorted(scope_stats.get("per_scope", {}).items(), key=lambda kv: -kv[1]["ns"])}
    last_loss = _finite_from_any_rank(last_loss, world)
    first_or_last = state.pre_process or state.post_process
    h2d = sum(v.numel() * v.element_size() for b in pin_batches for v in b.values()) if first_or_last else 0
    d2h = 4 if state.post_process else 0       # the loss scalar is read back on the last pipeline stage every step
    t = torch.tensor([ms, ms_e2e, ms_traced, exposed_ms, float(h2d), float(d2h)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e, ms_traced, exposed_ms = float(t[0]), float(t[1]), float(t[2]), float(t[3])
    h2d, d2h = int(t[4]), int(t[5])            # per-rank maxima: the stages that own the data / the loss
    tokens = global_batch * S * args.steps
    if rank == 0:
        par = (f"tp{tp}" + (f"pp{pp}" if pp > 1 else "") + (f"vpp{vpp}" if vpp else "") + (f"ep{ep}" if ep > 1 else "")
               + (f"dp{dp}" if dp > 1 else "") + ("+sp" if tp > 1 else "") + ("+fbd" if args.fbd else "")
               + (f"+dpp:{args.dpp}" if args.dpp else ""))
        out = {
            "metric": "GPT training tokens/sec (whole job, device-timed, max over ranks)",
            "value": tokens / (ms / 1e3), "unit": "tokens/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic tokens, random-init weights",
            "impl": "megatronapp_b200",
            "config": {"model": args.model, "layers": L, "hidden": H, "heads": A, "ffn": F, "vocab_padded": vocab,
                       "global_batch": global_batch, "micro_batch": mbs, "seq_len": S, "parallelism": par,
                       "optimizer": "Adam fp32 master + fp32 main grads, clip 1.0" + (", bf16 moments" if bf16_moments else ""),
                       "l2_policy": "working set (>=13 GB weights + grads per step) >> 126 MB L2; no explicit flush",
                       "attention_impl": "native" if ops.native_attention_available() else "sdpa-library",
                       "fused_tp_comm": (tp > 1 and not args.no_fused_tp_comm),
                       "fused_gemm_rs": (tp > 1 and not args.no_fused_tp_comm and not args.no_fused_rs),
                       "recompute_layers": recompute_layers, "adam_moments": "bf16" if bf16_moments else "fp32",
                       "dp_grad_sync": (None if dp == 1 else ("fused-kernel" if (args.fused_grad_rs and not moe) else "nccl")
                                        + ("" if args.no_dp_overlap else "+overlap"))},
            "clocks": clocks,
            "e2e": {"value": tokens / (ms_e2e / 1e3), "unit": "tokens/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "copies": "pinned-host microbatches -> device on the first/last pipeline stage every step; the "
                              "last stage reads the loss back (.item()) every step",
                    "final_loss": last_loss},
            "trace_overhead_pct": 100.0 * (ms_traced - ms) / ms,
            "traced_ms_per_step": ms_traced / args.steps,
            "exposed_comm_ms_per_step": exposed_ms,
            "gpu_launches": launches,
            "library_fallbacks": dict(ops.fallback_counts),
            "tflops_per_gpu": 6.0 * _param_count(L, H, F, vocab, arch, A) * tokens / (ms / 1e3) / 1e12 / args.gpus,
        }
        out["rank0_scope_ms_per_step"] = scope_ms
        if detector is not None:
            out["detector"] = detector
        if slow is not None:
            out["slow_rank_inject"] = {"rank": slow[0], "cycles_per_injection": slow[1], "injections_per_layer": 4}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


COMM_SCOPES = {"_reduce", "_gather_along_first_dim", "_reduce_scatter_along_first_dim", "_gather_along_last_dim",
               "_reduce_scatter_along_last_dim", "all_to_all", "allreduce", "reduce", "grad-reduce-scatter",
               "param-all-gather", "send-forward", "recv-forward", "send-backward", "recv-backward",
               "send-forward-recv-backward", "send-backward-recv-forward", "fbd-send", "fbd-recv"}


def _digest_scopes(records):
    """Per-scope device time and the exposed communication time of this rank from MegaScan B/E records.
    Exposed = time the compute stream spends inside (outermost) NCCL collective / p2p scopes; the fused
    AG->GEMM / GEMM->RS kernels carry their transfers inside the GEMM and are listed under their own scope names."""
    per, stack, comm_depth, exposed = {}, [], 0, 0
    for r in records:
        ph, name = r.get("ph"), r.get("name")
        if name == "iteration":
            continue
        if ph == "B":
            stack.append((name, r["rel_ts"]))
            if name in COMM_SCOPES:
                comm_depth += 1
        elif ph == "E":
            for k in range(len(stack) - 1, -1, -1):
                if stack[k][0] == name:
                    _, t0 = stack.pop(k)
                    d = per.setdefault(name, {"ns": 0, "calls": 0})
                    d["ns"] += r["rel_ts"] - t0
                    d["calls"] += 1
                    if name in COMM_SCOPES:
                        comm_depth -= 1
                        if comm_depth == 0:
                            exposed += r["rel_ts"] - t0
                    break
    return {"per_scope": per, "exposed_comm_ns": exposed}


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


def _param_count(L, H, F, V, arch=None, heads=None):
    """ACTIVE matmul parameters per token (MoE: top-k experts), for the model-FLOPs estimate 6 * N * tokens."""
    arch = arch or {}
    kv = arch.get("kv_groups") or heads
    hd = H // heads if heads else H
    attn = H * H + 2 * H * (kv * hd if heads else H) + H * H
    mlp = (3 if arch.get("swiglu") else 2) * H * F * (arch.get("moe_topk", 1) if arch.get("num_experts") else 1)
    return L * (attn + mlp) + V * H


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
    ap.add_argument("--layout", default=None, help="e.g. tp2ep4, tp1pp4dp2vpp2, tp8 (overrides the per-N default)")
    ap.add_argument("--vpp", type=int, default=None, help="virtual pipeline stages per rank (interleaved 1F1B)")
    ap.add_argument("--global-batch", type=int, default=None, help="sequences per step (default 8 per GPU)")
    ap.add_argument("--seq-len", type=int, default=None)
    ap.add_argument("--fbd", action="store_true", help="MegaFBD: forward and backward of a stage on separate GPUs")
    ap.add_argument("--dpp", default=None, choices=["static", "bfc", "dfc", "auto"], help="MegaDPP dynamic pipeline schedule")
    ap.add_argument("--fp8", action="store_true", help="fp8 (hybrid e4m3/e5m2) linears")
    ap.add_argument("--fp8-recipe", default="delayed", choices=["delayed", "tensorwise", "mxfp8"])
    ap.add_argument("--slow-rank-inject", default=None, help="RANK:CYCLES software slow-rank injector (MegaScan demo)")
    ap.add_argument("--no-dp-overlap", action="store_true", help="dp>1: reduce-scatter the gradients after backward instead of per bucket")
    ap.add_argument("--fused-grad-rs", action="store_true",
                    help="dp>1: sm_100a bucket reduce-scatter (scale + cast + RS over NVLink peer stores) instead of NCCL")
    ap.add_argument("--bf16-adam-moments", action="store_true", help="precision-aware optimizer: exp_avg / exp_avg_sq in bf16")
    ap.add_argument("--recompute-layers", type=int, default=0,
                    help="full activation recompute of the first N layers per stage (memory-bound configs, e.g. Mixtral 8x7B on 8 GPUs)")
    ap.add_argument("--detect", action="store_true", help="run the MegaScan detector on the traced steps (rank 0) and report it")
    ap.add_argument("--moe-baseline", action="store_true",
                    help="MoE in-tree baseline: NCCL all-to-all dispatcher + one module per expert (reference algorithm)")
    ap.add_argument("--no-fused-tp-comm", action="store_true", help="NCCL collectives instead of fused AG/RS GEMMs")
    ap.add_argument("--trace-scopes", action="store_true", help="(kept for compatibility; scopes are always reported)")
    ap.add_argument("--no-fused-rs", action="store_true", help="GEMM + NCCL reduce-scatter instead of the fused GEMM->RS kernel")
    ap.add_argument("--fused-rs", action="store_true", help="(default now) fuse GEMM->reduce-scatter")
    args = ap.parse_args()
    if args.no_fused_rs:
        os.environ["B200_FUSED_RS"] = "0"
    if args.impl == "reference":
        from baseline.reference_arm import run_reference

        run_reference(args, MODELS, layout_for, pad_vocab)
    else:
        run_own(args)


if __name__ == "__main__":
    main()
