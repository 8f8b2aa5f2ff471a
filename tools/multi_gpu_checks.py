"""Multi-GPU checks of the fused compute+collective kernels (run under torchrun, N >= 2):

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tools/multi_gpu_checks.py --out gpurun_out/multi.json

Numerics: fused kernel vs (NCCL collective + fp32 matmul); timing: device events, max over
ranks, against the unfused cuBLAS + NCCL sequence on the same tensors.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def rel_err(a, b):
    """element-wise: max |a - b| / (|b| + rms(b))  (same criterion as tools/gpu_checks.py)"""
    a, b = a.float(), b.float()
    rms = b.pow(2).mean().sqrt().clamp_min(1e-12)
    return float(((a - b).abs() / (b.abs() + rms)).max())


def roofline_ms(flops, nvlink_bytes):
    """B200_PROFILING.md: slower of FLOPs / measured GEMM peak and NVLink bytes / 770 GB/s per direction."""
    peak = 1671.6e12
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"] * 1e12
    except Exception:
        pass
    return max(flops / peak, nvlink_bytes / 770e9) * 1e3


def timed(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def dp_and_embed_checks(arena, fc, group, rank, world, dt, log, ops):
    # ---------------- fused bucket reduce-scatter (scale + cast + RS): a default-sized DP bucket (40 M fp32 elements)
    n = 40 * 1024 * 1024
    torch.manual_seed(400 + rank)
    g = torch.randn(n, device="cuda")
    ref = torch.empty(n // world, device="cuda")
    dist.reduce_scatter_tensor(ref, g / world, group=group)
    stage = arena.region("rs_stage", n * 4)
    flags = arena.region("rs_flags", 128 * 8 * 4)
    outs = torch.empty(n // world, device="cuda")

    def nccl():
        g2 = g * (1.0 / world)
        dist.reduce_scatter_tensor(ref, g2, group=group)

    tn_ = timed(nccl, 10, 3)
    for wire, wname in ((2, "fp32"), (0, "bf16")):
        ep = arena.next_epoch(flags)
        ops.ext().bucket_reduce_scatter(g, outs, arena.ptrs(stage), arena.ptrs(flags), rank, world, 1.0 / world, ep, wire, 32)
        torch.cuda.synchronize()
        log(f"bucket_rs_{wname}_err", rel_err(outs, ref))
        row = {"mul+nccl": tn_}
        for ctas in (16, 32, 64, 128):
            def fused(ctas=ctas):
                e = arena.next_epoch(flags)
                ops.ext().bucket_reduce_scatter(g, outs, arena.ptrs(stage), arena.ptrs(flags), rank, world, 1.0 / world, e, wire, ctas)

            t = timed(fused, 10, 3)
            # bytes this rank puts on the wire / time, and the fp32-equivalent ("algorithm") rate NCCL is quoted in
            row[f"fused_{ctas}cta"] = t
            row[f"wire_GBs_{ctas}cta"] = n * (4 if wire == 2 else 2) * (world - 1) / world / t / 1e6
            row[f"fp32_equiv_GBs_{ctas}cta"] = n * 4 * (world - 1) / world / t / 1e6
        row["nccl_fp32_GBs"] = n * 4 * (world - 1) / world / tn_ / 1e6
        log(f"bucket_rs_{wname}_ms", row)

    # ---------------- K14: vocab-parallel embedding fused with the sequence-parallel scatter
    import torch.nn.functional as F_

    bb, ss, hh, Vv = world, 2048, 4096, 50432 // world * world
    torch.manual_seed(11)
    toks = torch.randint(0, Vv, (bb, ss), device="cuda")
    torch.manual_seed(500 + rank)
    wemb = torch.randn(Vv // world, hh, device="cuda", dtype=dt)
    v0 = rank * (Vv // world)

    def emb_ref():
        m = (toks < v0) | (toks >= v0 + Vv // world)
        o = F_.embedding((toks - v0).masked_fill(m, 0), wemb).masked_fill(m.unsqueeze(-1), 0.0).transpose(0, 1).contiguous()
        r = torch.empty(ss // world, bb, hh, device="cuda", dtype=dt)
        dist.reduce_scatter_tensor(r, o, group=group)
        return r

    if fc.embed_fits(bb * ss, hh):
        got, want = fc.embed(toks, wemb, v0), emb_ref()
        torch.cuda.synchronize()
        log("embed_scatter_exact", bool(torch.equal(got, want)))
        log("embed_scatter_ms", {"fused": timed(lambda: fc.embed(toks, wemb, v0)), "embedding+mask+nccl_rs": timed(emb_ref)})
    else:
        log("embed_scatter_exact", "arena too small")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "multi.json"))
    ap.add_argument("--skip-model", action="store_true")
    ap.add_argument("--skip-moe", action="store_true")
    ap.add_argument("--skip-kernels", action="store_true")
    ap.add_argument("--dp-only", action="store_true", help="only the bucket reduce-scatter sweep and the K14 embedding scatter")
    a = ap.parse_args()
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    from megatronapp_b200 import ops
    from megatronapp_b200.comm.arena import SymmetricArena, TPFusedComm
    from megatronapp_b200.parallel import state as ps

    res = {"world": world}
    ps.initialize_model_parallel(tensor_model_parallel_size=world)
    group = ps.get_tensor_model_parallel_group()
    dt = torch.bfloat16

    def log(k, v):
        res[k] = v
        if rank == 0:
            print(k, v, flush=True)

    try:
        if a.skip_kernels:
            raise KeyboardInterrupt
        if a.dp_only:
            Mt = 2048 * world
            arena = SymmetricArena(group, (520 << 20))
            fc = TPFusedComm(arena, Mt, 4096, 4096)
            arena.barrier()
            dp_and_embed_checks(arena, fc, group, rank, world, dt, log, ops)
            arena.barrier()
            torch.cuda.synchronize()
            raise KeyboardInterrupt
        Mt = 2048 * world          # bench.py: micro-batch = TP degree, seq 2048
        shapes = [(Mt, 4096, 16384 // world, "fc1"), (Mt, 4096, 12288 // world, "qkv")]
        arena = SymmetricArena(group, 2 * (2 * Mt * 16384 * 2) + (520 << 20))
        fc = TPFusedComm(arena, Mt, 16384, 16384)
        arena.barrier()
        torch.cuda.synchronize()
        log("arena_ok", True)
        variants = (("2cta", 0), ("1cta", 256))
        # ---------------- AG -> GEMM (column parallel forward, bias + GELU + pre-activation in the epilogue)
        for (M, K, N, name) in shapes:
            torch.manual_seed(100 + rank)
            x_local = torch.randn(M // world, K, device="cuda", dtype=dt) * 0.5
            torch.manual_seed(7)
            w = torch.randn(N, K, device="cuda", dtype=dt) * 0.05
            bias = torch.randn(N, device="cuda", dtype=dt)
            full = torch.empty(M, K, device="cuda", dtype=dt)
            dist.all_gather_into_tensor(full, x_local, group=group)
            ref = full.float() @ w.float().t() + bias.float()
            outb = torch.empty(M, N, device="cuda", dtype=dt)
            preb = torch.empty(M, N, device="cuda", dtype=dt)
            times = {}
            for vname, bn in variants:
                fc.block_n = bn
                out, gathered = fc.ag_gemm(x_local, w, bias)
                torch.cuda.synchronize()
                log(f"ag_gemm_{name}_{vname}_err", rel_err(out, ref))
                log(f"ag_gemm_{name}_{vname}_gather_err", rel_err(gathered, full))
                out, _ = fc.ag_gemm(x_local, w, bias, act=1, out_pre=preb)
                torch.cuda.synchronize()
                log(f"ag_gemm_{name}_{vname}_gelu_err", max(rel_err(preb, ref), rel_err(out, torch.nn.functional.gelu(ref, approximate="tanh"))))
                sweep = {}
                for ncc in ((8, 16) if world > 2 else (4, 8)):
                    fc.n_comm_ctas = ncc
                    sweep[ncc] = timed(lambda: fc.ag_gemm(x_local, w, bias, out=outb), 10, 3)
                best = min(sweep, key=sweep.get)
                fc.n_comm_ctas = best
                times[f"fused_{vname}"] = timed(lambda: fc.ag_gemm(x_local, w, bias, out=outb))
                times[f"fused_{vname}_comm_ctas"] = sweep

            def unfused():
                dist.all_gather_into_tensor(full, x_local, group=group)
                ops.gemm(full, w, out=outb, bias=bias)

            def cublas_unfused():
                dist.all_gather_into_tensor(full, x_local, group=group)
                torch.addmm(bias, full, w.t(), out=outb)

            times["nccl+b200gemm"] = timed(unfused)
            times["nccl+cublas"] = timed(cublas_unfused)
            times["gemm_only"] = timed(lambda: ops.gemm(full, w, out=outb, bias=bias))
            best_f = min(times["fused_2cta"], times["fused_1cta"])
            roof = roofline_ms(2.0 * M * N * K, (M // world) * K * 2 * (world - 1))
            times.update(roofline_ms=roof, frac_of_roofline=roof / best_f, fused_over_cublas_nccl=best_f / times["nccl+cublas"],
                         tflops_fused=2.0 * M * N * K / best_f / 1e9)
            log(f"ag_gemm_{name}_ms", times)
        # ---------------- GEMM -> RS (row parallel forward)
        for (M, N, K, name) in [(Mt, 4096, 16384 // world, "fc2"), (Mt, 4096, 4096 // world, "proj")]:
            torch.manual_seed(200 + rank)
            x = torch.randn(M, K, device="cuda", dtype=dt) * 0.5
            w = torch.randn(N, K, device="cuda", dtype=dt) * 0.05
            bias = torch.randn(N, device="cuda", dtype=dt)
            part = (x.float() @ w.float().t())
            ref = torch.empty(M // world, N, device="cuda", dtype=torch.float32)
            dist.reduce_scatter_tensor(ref, part, group=group)
            pb = torch.empty(M, N, device="cuda", dtype=dt)
            ob = torch.empty(M // world, N, device="cuda", dtype=dt)
            times = {}
            for vname, bn in variants:
                fc.block_n = bn
                out = fc.gemm_rs(x, w)
                torch.cuda.synchronize()
                log(f"gemm_rs_{name}_{vname}_err", rel_err(out, ref))
                out = fc.gemm_rs(x, w, bias=bias)
                torch.cuda.synchronize()
                log(f"gemm_rs_{name}_{vname}_bias_err", rel_err(out, ref + bias.float()))
                o1 = fc.gemm_rs(x, w).clone()
                o2 = fc.gemm_rs(x, w)
                log(f"gemm_rs_{name}_{vname}_deterministic", bool(torch.equal(o1, o2)))
                times[f"fused_{vname}"] = timed(lambda: fc.gemm_rs(x, w))

            def unfused():
                ops.gemm(x, w, out=pb)
                dist.reduce_scatter_tensor(ob, pb, group=group)

            def cublas_unfused():
                torch.matmul(x, w.t(), out=pb)
                dist.reduce_scatter_tensor(ob, pb, group=group)

            times["b200gemm+nccl"] = timed(unfused)
            times["cublas+nccl"] = timed(cublas_unfused)
            times["gemm_only"] = timed(lambda: ops.gemm(x, w, out=pb))
            best_f = min(times["fused_2cta"], times["fused_1cta"])
            roof = roofline_ms(2.0 * M * N * K, (M // world) * N * 2 * (world - 1))
            times.update(roofline_ms=roof, frac_of_roofline=roof / best_f, fused_over_cublas_nccl=best_f / times["cublas+nccl"],
                         tflops_fused=2.0 * M * N * K / best_f / 1e9)
            log(f"gemm_rs_{name}_ms", times)
        # dgrad duals (B MN-major)
        M, K, N = Mt, 4096, 16384 // world
        torch.manual_seed(300 + rank)
        dy = torch.randn(M, N, device="cuda", dtype=dt) * 0.5
        w = torch.randn(N, K, device="cuda", dtype=dt) * 0.05
        part = dy.float() @ w.float()
        ref = torch.empty(M // world, K, device="cuda", dtype=torch.float32)
        dist.reduce_scatter_tensor(ref, part, group=group)
        dyl = torch.randn(M // world, K, device="cuda", dtype=dt) * 0.5   # row-parallel dgrad: AG(dy) @ W[K_out, N_in]
        w2 = torch.randn(K, N, device="cuda", dtype=dt) * 0.05
        fulldy = torch.empty(M, K, device="cuda", dtype=dt)
        dist.all_gather_into_tensor(fulldy, dyl, group=group)
        ref2 = fulldy.float() @ w2.float()
        pbd = torch.empty(M, K, device="cuda", dtype=dt)
        obd = torch.empty(M // world, K, device="cuda", dtype=dt)
        for vname, bn in variants:
            fc.block_n = bn
            out = fc.gemm_rs(dy, w, b_mn=True)
            log(f"gemm_rs_dgrad_{vname}_err", rel_err(out, ref))
            out2, _ = fc.ag_gemm(dyl, w2, None, b_mn=True)
            log(f"ag_gemm_dgrad_{vname}_err", rel_err(out2, ref2))
            log(f"gemm_rs_dgrad_{vname}_ms", timed(lambda: fc.gemm_rs(dy, w, b_mn=True)))

        def dgrad_unfused():
            torch.matmul(dy, w, out=pbd)
            dist.reduce_scatter_tensor(obd, pbd, group=group)

        log("gemm_rs_dgrad_cublas+nccl_ms", timed(dgrad_unfused))
        fc.block_n = 0
        # ---------------- AG -> wgrad (column-parallel weight gradient: all-gather fused along the reduction axis)
        for (tok, n_out, k_in, name) in [(Mt, 16384 // world, 4096, "fc1"), (Mt, 12288 // world, 4096, "qkv")]:
            torch.manual_seed(500 + rank)
            x_l = torch.randn(tok // world, k_in, device="cuda", dtype=dt) * 0.5
            dyw = torch.randn(tok, n_out, device="cuda", dtype=dt) * 0.1
            xf = torch.empty(tok, k_in, device="cuda", dtype=dt)
            dist.all_gather_into_tensor(xf, x_l, group=group)
            refw = dyw.float().t() @ xf.float()
            g32 = torch.zeros(n_out, k_in, device="cuda")
            fc.ag_wgrad(dyw, x_l, g32, accumulate=False)
            torch.cuda.synchronize()
            log(f"ag_wgrad_{name}_err", rel_err(g32, refw))
            fc.ag_wgrad(dyw, x_l, g32, accumulate=True)
            torch.cuda.synchronize()
            log(f"ag_wgrad_{name}_accum_err", rel_err(g32, 2 * refw))

            def wg_unfused():
                dist.all_gather_into_tensor(xf, x_l, group=group)
                ops.gemm(dyw, xf, out=g32, a_mn=True, b_mn=True, epilogue=1)

            tw = {"fused": timed(lambda: fc.ag_wgrad(dyw, x_l, g32, accumulate=True)), "nccl+b200gemm": timed(wg_unfused),
                  "gemm_only": timed(lambda: ops.gemm(dyw, xf, out=g32, a_mn=True, b_mn=True, epilogue=1))}
            roof = roofline_ms(2.0 * tok * n_out * k_in, (tok // world) * k_in * 2 * (world - 1))
            tw.update(roofline_ms=roof, frac_of_roofline=roof / tw["fused"], tflops_fused=2.0 * tok * n_out * k_in / tw["fused"] / 1e9)
            log(f"ag_wgrad_{name}_ms", tw)

        dp_and_embed_checks(arena, fc, group, rank, world, dt, log, ops)
        arena.barrier()
        torch.cuda.synchronize()
    except KeyboardInterrupt:
        pass
    except Exception as e:  # noqa: BLE001
        import traceback

        log("EXCEPTION_kernels", traceback.format_exc()[-1500:])
    # ---------------- model parity: fused TP path vs NCCL path (same weights, same data)
    if not a.skip_model:
        try:
            from megatronapp_b200.bench_model import build_gpt_for_bench, synthetic_batch, train_step
            from megatronapp_b200.parallel import tp_layers

            ps.destroy_model_parallel()
            st = build_gpt_for_bench(2, 1024, 8, 1024, 8192, 1, ffn=4096, tp=world, sequence_parallel=True, device="cuda")
            b = synthetic_batch(st, "cuda")
            sd = {k: v.clone() for k, v in st.model[0].module.module.state_dict().items()}
            fused_losses = [train_step(st, [b, b]) for _ in range(3)]
            log("fused_calls", dict(tp_layers.get_tp_fused_comm().calls) if tp_layers.get_tp_fused_comm() else None)
            tp_layers.disable_tp_fused_comm()
            ps.destroy_model_parallel()
            st2 = build_gpt_for_bench(2, 1024, 8, 1024, 8192, 1, ffn=4096, tp=world, sequence_parallel=True,
                                      device="cuda", fused_tp_comm=False)
            st2.model[0].module.module.load_state_dict(sd)
            st2.optimizer.reload_model_params()
            nccl_losses = [train_step(st2, [b, b]) for _ in range(3)]
            log("tp_model_losses", {"fused": fused_losses, "nccl": nccl_losses})
        except Exception:  # noqa: BLE001
            import traceback

            log("EXCEPTION_model", traceback.format_exc()[-1500:])
    # ---------------- expert parallel MoE layer: arena dispatcher + grouped GEMMs vs NCCL all-to-all + per-expert modules
    if not a.skip_moe:
        try:
            import torch.nn.functional as F
            from megatronapp_b200.comm.arena import destroy_arenas
            from megatronapp_b200.config.transformer_config import TransformerConfig
            from megatronapp_b200.models.gpt.gpt_layer_specs import get_mlp_module_spec
            from megatronapp_b200.parallel.random import model_parallel_cuda_manual_seed
            from megatronapp_b200.transformer.spec_utils import build_module

            ps.destroy_model_parallel()
            ps.initialize_model_parallel(1, 1, expert_model_parallel_size=world)
            model_parallel_cuda_manual_seed(123)
            E = 8 if 8 % world == 0 else 2 * world

            def build(kind, grouped, H, FF):
                cfg = TransformerConfig(num_layers=1, hidden_size=H, num_attention_heads=8, ffn_hidden_size=FF, num_moe_experts=E,
                                        moe_router_topk=2, expert_model_parallel_size=world, moe_token_dispatcher_type=kind,
                                        add_bias_linear=False, activation_func=F.silu, gated_linear_unit=True,
                                        moe_grouped_gemm=grouped, bf16=True, params_dtype=torch.bfloat16, moe_aux_loss_coeff=0.0)
                layer = build_module(get_mlp_module_spec(False, E, grouped), config=cfg).cuda()
                with torch.no_grad():
                    torch.manual_seed(11)
                    layer.router.weight.copy_(torch.randn_like(layer.router.weight.float()) * 0.3)
                    for i, e in enumerate(layer.local_expert_indices):
                        g = torch.Generator().manual_seed(1000 + e)
                        w1 = (torch.randn(2 * FF, H, generator=g) * 0.03).bfloat16().cuda()
                        w2 = (torch.randn(H, FF, generator=g) * 0.03).bfloat16().cuda()
                        if grouped:
                            layer.experts.weight1[i * 2 * FF:(i + 1) * 2 * FF].copy_(w1)
                            layer.experts.weight2[i * H:(i + 1) * H].copy_(w2)
                        else:
                            layer.experts.local_experts[i].linear_fc1.weight.copy_(w1)
                            layer.experts.local_experts[i].linear_fc2.weight.copy_(w2)
                return layer

            def run(layer, x, go):
                x = x.clone().requires_grad_(True)
                for p_ in layer.parameters():
                    p_.grad = None
                out, _ = layer(x)
                (out.float() * go).sum().backward()
                return out.detach(), x.grad.detach()

            H, FF, T = 1024, 2048, 1024
            gx = torch.Generator().manual_seed(77 + rank)
            x = (torch.randn(T, 2, H, generator=gx) * 0.5).bfloat16().cuda()
            go = torch.randn(T, 2, H, generator=torch.Generator().manual_seed(5 + rank)).cuda()
            ref_layer = build("alltoall", False, H, FF)
            o_ref, dx_ref = run(ref_layer, x, go)
            w1_ref = torch.cat([ex.linear_fc1.weight.grad.float() for ex in ref_layer.experts.local_experts])
            w2_ref = torch.cat([ex.linear_fc2.weight.grad.float() for ex in ref_layer.experts.local_experts])
            dr_ref = ref_layer.router.weight.grad.float().clone()
            fl = build("flex", True, H, FF)
            o_f, dx_f = run(fl, x, go)
            log("moe_ep_dispatcher", type(fl.token_dispatcher).__name__)
            log("moe_ep_err", {"out": rel_err(o_f, o_ref), "dx": rel_err(dx_f, dx_ref),
                               "drouter": rel_err(fl.router.weight.grad.float(), dr_ref),
                               "dw1": rel_err(fl.experts.weight1.grad.float(), w1_ref),
                               "dw2": rel_err(fl.experts.weight2.grad.float(), w2_ref),
                               "overflow": bool(fl.token_dispatcher.check_overflow())})
            o_f2, dx_f2 = run(fl, x, go)
            log("moe_ep_deterministic", bool(torch.equal(o_f, o_f2) and torch.equal(dx_f, dx_f2)))
            del ref_layer, fl
            destroy_arenas()
            # timing at Mixtral-8x7B layer dimensions (h 4096, ffn 14336, 8 experts, top-2), 4096 tokens per rank
            H, FF, T = 4096, 14336, 2048
            x = (torch.randn(T, 2, H, generator=gx) * 0.5).bfloat16().cuda()
            go = torch.randn(T, 2, H, generator=torch.Generator().manual_seed(5 + rank)).cuda()
            tm = {}
            for name, kind, grouped in (("arena+grouped", "flex", True), ("nccl_a2a+grouped_loop", "alltoall", True),
                                        ("nccl_a2a+sequential", "alltoall", False)):
                layer = build(kind, grouped, H, FF)
                tm[name] = timed(lambda: run(layer, x, go), 8, 3)
                del layer
                torch.cuda.empty_cache()
            flops = 3 * 2.0 * (T * 2 * 2) * H * (3 * FF)          # fwd + bwd, top-2, gate/up/down
            tm["tflops_per_gpu_arena"] = flops / tm["arena+grouped"] / 1e9
            log("moe_ep_layer_fwd_bwd_ms", tm)
        except Exception:  # noqa: BLE001
            import traceback

            log("EXCEPTION_moe", traceback.format_exc()[-2500:])
    if rank == 0:
        os.makedirs(os.path.dirname(a.out), exist_ok=True)
        json.dump(res, open(a.out, "w"), indent=1)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
