"""Small-shape exercise of the flag-protocol / peer-memory kernels, meant to run under compute-sanitizer:

    python -m torch.distributed.run --no-python --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29599 \\
        compute-sanitizer --tool synccheck --log-file gpurun_out/synccheck_%p.log python tools/sanitize_comm.py

(fused AG->GEMM, GEMM->RS, bucket reduce-scatter, embedding scatter, device barrier; with WORLD_SIZE=1 the single-GPU
shared-memory kernels instead: paged attention, MX quantiser, norms, MoE layout / gather kernels).  Prints OK lines so the
log shows every kernel really ran under the tool."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def multi():
    from megatronapp_b200 import ops
    from megatronapp_b200.comm.arena import SymmetricArena, TPFusedComm
    from megatronapp_b200.parallel import state as ps

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    ps.initialize_model_parallel(tensor_model_parallel_size=world)
    group = ps.get_tensor_model_parallel_group()
    M, K, N = 512 * world, 1024, 1024
    arena = SymmetricArena(group, 64 << 20)
    fc = TPFusedComm(arena, M, max(K, N), max(K, N))
    arena.barrier()
    dt = torch.bfloat16
    torch.manual_seed(rank)
    x = torch.randn(M // world, K, device="cuda", dtype=dt)
    w = torch.randn(N, K, device="cuda", dtype=dt) * 0.05
    for _ in range(2):
        out, _ = fc.ag_gemm(x, w)
    full = torch.empty(M, K, device="cuda", dtype=dt)
    dist.all_gather_into_tensor(full, x, group=group)
    err = ((out.float() - full.float() @ w.float().t()).abs().max() / (full.float() @ w.float().t()).abs().max()).item()
    print(f"[{rank}] OK ag_gemm err={err:.4f}", flush=True)
    xr = torch.randn(M, K, device="cuda", dtype=dt)
    for _ in range(2):
        o = fc.gemm_rs(xr, w)
    print(f"[{rank}] OK gemm_rs {tuple(o.shape)}", flush=True)
    n = 1 << 16
    g = torch.randn(n, device="cuda")
    stage, flags = arena.region("rs_stage", n * 4), arena.region("rs_flags", 128 * 8 * 4)
    outs = torch.empty(n // world, device="cuda")
    for wire in (2, 0):
        ops.ext().bucket_reduce_scatter(g, outs, arena.ptrs(stage), arena.ptrs(flags), rank, world, 1.0 / world,
                                        arena.next_epoch(flags), wire, 8)
    print(f"[{rank}] OK bucket_rs", flush=True)
    toks = torch.randint(0, 1024, (2, 128 * world), device="cuda")
    wemb = torch.randn(1024 // world, 1024, device="cuda", dtype=dt)
    if fc.embed_fits(toks.numel(), 1024):
        e = fc.embed(toks, wemb, rank * (1024 // world))
        print(f"[{rank}] OK embed_scatter {tuple(e.shape)}", flush=True)
    arena.barrier()
    torch.cuda.synchronize()
    dist.barrier()
    print(f"[{rank}] DONE", flush=True)
    dist.destroy_process_group()


def single():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
    import gpu_checks as gc

    torch.cuda.set_device(0)
    args, ref = gc._paged_case([1, 3, 1], [40, 0, 130], 8, 2, 128, 16, torch.bfloat16)
    from megatronapp_b200 import ops

    for splits in (1, 2):
        got = ops.ext().paged_attention(*args, splits)
    print("OK paged_attention", float((got.float() - ref).abs().max()), flush=True)
    from megatronapp_b200.ops import fp8

    x = torch.randn(256, 512, device="cuda", dtype=torch.bfloat16)
    q = fp8.quantize_mx_native(x)
    print("OK mx_quantize", type(q).__name__, flush=True)
    h = torch.randn(64, 1024, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    wln = torch.ones(1024, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    bln = torch.zeros(1024, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    ops.layer_norm(h, wln, bln, 1e-5).float().sum().backward()
    print("OK layer_norm fwd/bwd", flush=True)
    r = gc.check_moe_kernels()
    print("OK moe kernels", max(v for v in r.values() if isinstance(v, (int, float))), flush=True)
    torch.cuda.synchronize()
    print("DONE", flush=True)


if __name__ == "__main__":
    multi() if int(os.environ.get("WORLD_SIZE", "1")) > 1 else single()
