"""Pipeline point-to-point micro-benchmark: DPP NVLink mailbox vs NCCL send/recv vs the arena
peer copy, reported as GB/s against NVLink 5 (900 GB/s per direction).

Parity: reference ``profiling/shm_benchmark.cpp`` + ``shm_benchmark_test.py`` +
``baseline_benchmark.py`` (10 MB fp16 tensor through shm+semaphores vs ``dist.send/recv``).

    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 tools/bench_p2p.py [--mbytes 10,64,256]

Times are CUDA events on the sender's stream (receiver polls / waits), max over the two ranks.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist

    from megatronapp_b200 import ops

    ap = argparse.ArgumentParser()
    ap.add_argument("--mbytes", default="10,64,256")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    assert world == 2, "run with exactly two ranks"
    ext = ops.ext()
    peer = 1 - rank
    res = {}

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / a.iters], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    for mb in [int(x) for x in a.mbytes.split(",")]:
        n = mb * 1024 * 1024 // 2
        x = torch.randn(n, device="cuda", dtype=torch.float16)
        y = torch.empty_like(x)

        def nccl():
            if rank == 0:
                dist.send(x, 1)
            else:
                dist.recv(y, 0)

        ms = timed(nccl)
        res[f"nccl_sendrecv_{mb}MB_GBs"] = round(mb / 1024 / (ms / 1e3), 1)
        # mailbox: one slot, epoch-stamped doorbell in POSIX shm, payload by NVLink peer copy
        name = f"/b200p2p_{os.environ.get('MASTER_PORT', '0')}_{rank}"
        box, handle = ext.mailbox_create(1, n * 2, name)
        info = [None, None]
        dist.all_gather_object(info, (handle, name))
        ext.mailbox_connect(box, info[peer][0], info[peer][1])
        dist.barrier()
        state = {"epoch": 0}

        def mailbox():
            state["epoch"] += 1
            if rank == 0:
                ext.mailbox_put(box, 0, x, state["epoch"])
            else:
                while not ext.mailbox_try_get(box, 0, state["epoch"]):
                    pass

        ms = timed(mailbox)
        res[f"mailbox_{mb}MB_GBs"] = round(mb / 1024 / (ms / 1e3), 1)
        if rank == 1 and mb <= 64:
            got = ext.mailbox_slot_tensor(box, 0, [n], torch.float16, torch.cuda.current_device())
            xs = [None, None]
        dist.barrier()
        ext.mailbox_destroy(box)
        # raw device-to-device copy on one GPU as the local HBM ceiling for context
        ms = timed(lambda: y.copy_(x))
        res[f"local_d2d_{mb}MB_GBs"] = round(mb / 1024 / (ms / 1e3), 1)
    res["nvlink5_per_direction_GBs"] = 900
    if rank == 0:
        print(json.dumps(res), flush=True)
        if a.out:
            json.dump(res, open(a.out, "w"), indent=1)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
