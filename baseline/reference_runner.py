"""Drives the unmodified reference (baseline/_ref: megatron.core 0.12.2 + the rest of the
reference's ``megatron`` package + its ``pretrain_gpt.py``) through its own public API:
``initialize_megatron`` -> ``setup_model_and_optimizer(model_provider)`` -> ``train_step(
forward_step, ...)`` with ``--transformer-impl local`` (the stock non-TE path: torch.matmul /
cuBLAS GEMMs, unfused attention, NCCL collectives).  Only the two RDMA C++ extensions that the
reference imports unconditionally (training.py:120) and that cannot be built without
libibverbs headers are replaced by empty stub modules; they are unused unless ``--use-dpp``.
"""
from __future__ import annotations

import json
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def main(args, MODELS, layout_for, pad_vocab):
    import torch

    sys.path.insert(0, REF)
    for n in ("shm_tensor_new_rdma", "shm_tensor_new_rdma_pre_alloc"):
        sys.modules.setdefault(n, types.ModuleType(n))
    tp, pp = layout_for(args.gpus)
    if args.tp:
        tp, pp = args.tp, args.gpus // args.tp
    if getattr(args, "layout", None):
        import re

        lay = dict(re.findall(r"(tp|pp)(\d+)", args.layout))
        tp, pp = int(lay.get("tp", 1)), int(lay.get("pp", 1))
    L, H, A, F, S, V = MODELS[args.model]
    if args.layers:
        L = args.layers
    dp = args.gpus // (tp * pp)
    if not args.micro_batch:
        args.micro_batch = tp          # same rule as the framework's own arm (bench.py)
    gbs = getattr(args, "global_batch", None) or args.microbatches_per_gpu * args.gpus      # weak scaling: 8 sequences per GPU
    n_mb = gbs // (args.micro_batch * dp)
    vocab = pad_vocab(V, tp)
    argv = ["reference", "--transformer-impl", "local", "--use-mcore-models",
            "--num-layers", str(L), "--hidden-size", str(H), "--num-attention-heads", str(A),
            "--ffn-hidden-size", str(F), "--seq-length", str(S), "--max-position-embeddings", str(S),
            "--micro-batch-size", str(args.micro_batch), "--global-batch-size", str(gbs),
            "--tensor-model-parallel-size", str(tp), "--pipeline-model-parallel-size", str(pp),
            "--bf16", "--lr", "1e-4", "--min-lr", "1e-5", "--train-iters", "1000000", "--lr-decay-iters", "1000000",
            "--weight-decay", "0.01", "--clip-grad", "1.0", "--adam-beta2", "0.95",
            "--attention-dropout", "0.0", "--hidden-dropout", "0.0",
            "--tokenizer-type", "NullTokenizer", "--vocab-size", str(vocab - 1), "--make-vocab-size-divisible-by", "1",
            "--mock-data", "--no-create-attention-mask-in-dataloader", "--no-gradient-accumulation-fusion",
            "--no-masked-softmax-fusion", "--no-persist-layer-norm", "--distributed-backend", "nccl", "--log-interval", "1000000",
            "--eval-iters", "0", "--no-save-optim", "--no-save-rng", "--ckpt-format", "torch"]
    if args.model == "llama3_8b":
        argv += ["--swiglu", "--position-embedding-type", "rope", "--normalization", "RMSNorm", "--group-query-attention",
                 "--num-query-groups", "8", "--disable-bias-linear", "--no-position-embedding"]
    if getattr(args, "seq_len", None):
        i = argv.index("--seq-length")
        argv[i + 1] = str(args.seq_len)
        argv[argv.index("--max-position-embeddings") + 1] = str(args.seq_len)
    import importlib.util

    # the stock local spec falls back to torch LayerNorm without Apex/TE, and that norm refuses
    # sequence parallelism (core/transformer/torch_norm.py:32) -> plain TP (all-reduce) then
    sp = tp > 1 and importlib.util.find_spec("apex") is not None
    if sp:
        argv += ["--sequence-parallel"]
    if dp > 1:
        argv += ["--use-distributed-optimizer"]
    sys.argv = argv
    os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "1")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    os.environ.setdefault("RANK", "0")
    os.environ.setdefault("WORLD_SIZE", "1")
    os.environ.setdefault("LOCAL_RANK", "0")
    import torch.distributed as dist
    from megatron.core import mpu
    from megatron.core.distributed import finalize_model_grads
    from megatron.core.enums import ModelType
    from megatron.core.utils import get_model_config
    from megatron.training import get_args, get_timers
    from megatron.training.initialize import initialize_megatron
    from megatron.training.training import setup_model_and_optimizer, train_step
    import pretrain_gpt as ref_gpt

    initialize_megatron(args_defaults={"tokenizer_type": "NullTokenizer"})
    margs = get_args()
    model, optimizer, sched = setup_model_and_optimizer(ref_gpt.model_provider, ModelType.encoder_or_decoder)
    config = get_model_config(model[0])
    config.grad_scale_func = optimizer.scale_loss
    config.timers = get_timers()
    config.finalize_model_grads_func = finalize_model_grads
    for m in model:
        m.train()
    margs.curr_iteration = 0
    dev = torch.device("cuda", torch.cuda.current_device())

    def make_batches(pinned):
        out = []
        for i in range(n_mb):
            g = torch.Generator().manual_seed(1000 + i)
            toks = torch.randint(1, vocab, (args.micro_batch, S + 1), generator=g, dtype=torch.int64)
            b = {"tokens": toks[:, :-1].contiguous(), "labels": toks[:, 1:].contiguous(),
                 "loss_mask": torch.ones(args.micro_batch, S, dtype=torch.float32),
                 "position_ids": torch.arange(S, dtype=torch.int64).unsqueeze(0).expand(args.micro_batch, -1).contiguous()}
            out.append({k: (v.pin_memory() if pinned else v.to(dev)) for k, v in b.items()})
        return out

    dev_b, pin_b = make_batches(False), make_batches(True)
    needs_data = mpu.get_tensor_model_parallel_rank() == 0 and (mpu.is_pipeline_first_stage() or mpu.is_pipeline_last_stage())

    def step(batches, sync_loss):
        it = iter(batches) if needs_data else None
        r = train_step(ref_gpt.forward_step, it, model, optimizer, sched, config)
        margs.curr_iteration += 1
        if sync_loss and mpu.is_pipeline_last_stage(ignore_virtual=True) and r[0]:
            v = r[0].get("lm loss")
            return float(v.item() if torch.is_tensor(v) else v)
        return None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    from bench import ClockSampler

    for _ in range(args.warmup):
        step(dev_b, False)
    barrier()
    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step(dev_b, False)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else {}
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    last = None
    for _ in range(args.steps):
        last = step(pin_b, True)
    f1.record()
    barrier()
    ms2 = f0.elapsed_time(f1)
    t = torch.tensor([ms, ms2], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms2 = float(t[0]), float(t[1])
    tokens = gbs * S * args.steps
    if rank == 0:
        par = f"tp{tp}" + (f"pp{pp}" if pp > 1 else "") + (f"dp{dp}" if dp > 1 else "") + ("+sp" if sp else ("+no-sp(torch LayerNorm lacks SP; no Apex/TE offline)" if tp > 1 else ""))
        h2d = sum(v.numel() * v.element_size() for b in pin_b for v in b.values())
        print(json.dumps({
            "impl": "reference", "metric": "GPT training tokens/sec (whole job, device-timed, max over ranks)",
            "value": tokens / (ms / 1e3), "unit": "tokens/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic tokens, random-init weights",
            "config": {"model": args.model, "layers": L, "hidden": H, "heads": A, "ffn": F, "vocab_padded": vocab,
                       "global_batch": gbs, "micro_batch": args.micro_batch, "seq_len": S, "parallelism": par,
                       "reference_path": "megatron.core 0.12.2 --transformer-impl local (cuBLAS + NCCL, unfused attention)",
                       "l2_policy": "working set >> 126 MB L2; no explicit flush"},
            "clocks": clocks,
            "e2e": {"value": tokens / (ms2 / 1e3), "unit": "tokens/s", "ms_per_step": ms2 / args.steps,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "final_loss": last},
            "gpu_launches": 0}), flush=True)
    if world > 1:
        dist.barrier()
