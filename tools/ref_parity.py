"""Loss-trajectory parity against the UNMODIFIED reference (baseline/_ref), same weights, same data.

    python tools/ref_parity.py --steps 20 [--out gpurun_out/ref_parity.json]                 # 1 GPU
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
        tools/ref_parity.py --tp 2                                                            # TP2 (+SP on our side)

Builds the reference GPT through its own API (``initialize_megatron`` -> ``setup_model_and_optimizer(pretrain_gpt.
model_provider)``, ``--transformer-impl local``), copies its (TP-local) weights into this framework's model -- parameter
names and TP sharding conventions are the same -- and then steps both with identical micro-batches, optimizer settings
(Adam, lr 1e-4, wd 0.01, clip 1.0, beta2 0.95) and no dropout.  Prints both loss curves and the per-step difference.
Also the tool that explains any gap between the two arms of ``bench.py``: both start from *different* random inits
there (own RNG consumption order), here they start from the same point.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.path.join(ROOT, "baseline", "_ref")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--tp", type=int, default=1)
    ap.add_argument("--layers", type=int, default=4)
    ap.add_argument("--hidden", type=int, default=1024)
    ap.add_argument("--heads", type=int, default=8)
    ap.add_argument("--ffn", type=int, default=4096)
    ap.add_argument("--seq", type=int, default=512)
    ap.add_argument("--vocab", type=int, default=8192)
    ap.add_argument("--micro-batch", type=int, default=4)
    ap.add_argument("--n-mb", type=int, default=2)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "ref_parity.json"))
    a = ap.parse_args()
    import torch

    if not os.path.isdir(os.path.join(REF, "megatron")):
        print(json.dumps({"unavailable": "baseline/_ref is not installed"}))
        return
    os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "1")
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29541")
    os.environ.setdefault("RANK", "0")
    os.environ.setdefault("WORLD_SIZE", "1")
    os.environ.setdefault("LOCAL_RANK", "0")
    world = int(os.environ["WORLD_SIZE"])
    rank = int(os.environ["RANK"])
    tp = a.tp
    assert world == tp, "launch with --nproc-per-node == --tp"
    gbs = a.micro_batch * a.n_mb
    sys.path.insert(0, REF)
    for n in ("shm_tensor_new_rdma", "shm_tensor_new_rdma_pre_alloc"):
        sys.modules.setdefault(n, types.ModuleType(n))
    sys.argv = ["reference", "--transformer-impl", "local", "--use-mcore-models",
                "--num-layers", str(a.layers), "--hidden-size", str(a.hidden), "--num-attention-heads", str(a.heads),
                "--ffn-hidden-size", str(a.ffn), "--seq-length", str(a.seq), "--max-position-embeddings", str(a.seq),
                "--micro-batch-size", str(a.micro_batch), "--global-batch-size", str(gbs),
                "--tensor-model-parallel-size", str(tp), "--pipeline-model-parallel-size", "1",
                "--bf16", "--lr", "1e-4", "--min-lr", "1e-5", "--train-iters", "1000000", "--lr-decay-iters", "1000000",
                "--weight-decay", "0.01", "--clip-grad", "1.0", "--adam-beta2", "0.95",
                "--attention-dropout", "0.0", "--hidden-dropout", "0.0",
                "--tokenizer-type", "NullTokenizer", "--vocab-size", str(a.vocab - 1), "--make-vocab-size-divisible-by", "1",
                "--mock-data", "--no-create-attention-mask-in-dataloader", "--no-gradient-accumulation-fusion",
                "--no-masked-softmax-fusion", "--no-persist-layer-norm", "--distributed-backend", "nccl",
                "--log-interval", "1000000", "--eval-iters", "0", "--no-save-optim", "--no-save-rng", "--ckpt-format", "torch"]
    from megatron.core import mpu
    from megatron.core.distributed import finalize_model_grads
    from megatron.core.enums import ModelType
    from megatron.core.utils import get_model_config
    from megatron.training import get_args, get_timers
    from megatron.training.initialize import initialize_megatron
    from megatron.training.training import setup_model_and_optimizer, train_step as ref_train_step
    import pretrain_gpt as ref_gpt

    initialize_megatron(args_defaults={"tokenizer_type": "NullTokenizer"})
    margs = get_args()
    rmodel, ropt, rsched = setup_model_and_optimizer(ref_gpt.model_provider, ModelType.encoder_or_decoder)
    rcfg = get_model_config(rmodel[0])
    rcfg.grad_scale_func = ropt.scale_loss
    rcfg.timers = get_timers()
    rcfg.finalize_model_grads_func = finalize_model_grads
    for m in rmodel:
        m.train()
    margs.curr_iteration = 0
    dev = torch.device("cuda", torch.cuda.current_device())

    # ---- this framework, same weights
    from megatronapp_b200.bench_model import build_gpt_for_bench, train_step

    st = build_gpt_for_bench(a.layers, a.hidden, a.heads, a.seq, a.vocab, a.micro_batch, ffn=a.ffn, tp=tp,
                             sequence_parallel=tp > 1, device=f"cuda:{dev.index}", lr=1e-4)
    ref_core = rmodel[0]
    while hasattr(ref_core, "module"):
        ref_core = ref_core.module
    ref_sd = {k: v.detach().clone() for k, v in ref_core.state_dict().items() if torch.is_tensor(v) and "_extra_state" not in k}
    ours = st.model[0].module.module
    own_sd = ours.state_dict()
    missing = [k for k in own_sd if k not in ref_sd]
    extra = [k for k in ref_sd if k not in own_sd]
    shape_bad = [k for k in own_sd if k in ref_sd and tuple(own_sd[k].shape) != tuple(ref_sd[k].shape)]
    assert not missing and not shape_bad, (missing[:5], extra[:5], shape_bad[:5])
    # both arms of bench.py start from their OWN random init: compare the init distributions parameter by parameter
    init_stats = {}
    for k, v in own_sd.items():
        if v.dim() >= 2 or k.endswith("position_embeddings.weight"):
            so, sr = float(v.float().std()), float(ref_sd[k].float().std())
            kk = ".".join(p for p in k.split(".") if not p.isdigit())
            init_stats.setdefault(kk, []).append(round(so / max(sr, 1e-12), 4))
    init_stats = {k: (min(v), max(v)) for k, v in init_stats.items()}
    with torch.no_grad():
        for k, v in own_sd.items():
            v.copy_(ref_sd[k].to(v.dtype))
    st.optimizer.reload_model_params()

    def make_batches():
        out = []
        for i in range(a.n_mb):
            g = torch.Generator().manual_seed(1000 + i)
            toks = torch.randint(1, a.vocab, (a.micro_batch, a.seq + 1), generator=g, dtype=torch.int64)
            out.append({"tokens": toks[:, :-1].contiguous().to(dev), "labels": toks[:, 1:].contiguous().to(dev),
                        "loss_mask": torch.ones(a.micro_batch, a.seq, dtype=torch.float32, device=dev),
                        "position_ids": torch.arange(a.seq, dtype=torch.int64, device=dev).unsqueeze(0).expand(a.micro_batch, -1).contiguous()})
        return out

    batches = make_batches()
    needs_data = mpu.get_tensor_model_parallel_rank() == 0
    ref_losses, own_losses = [], []
    for s in range(a.steps):
        it = iter(batches) if needs_data else None
        r = ref_train_step(ref_gpt.forward_step, it, rmodel, ropt, rsched, rcfg)
        margs.curr_iteration += 1
        v = r[0].get("lm loss") if r[0] else None
        ref_losses.append(float(v.item() if torch.is_tensor(v) else v) if v is not None else float("nan"))
        own_losses.append(train_step(st, batches, sync_loss=True))
    diffs = [abs(x - y) for x, y in zip(ref_losses, own_losses)]
    res = {"tp": tp, "steps": a.steps, "reference": ref_losses, "own": own_losses, "abs_diff": diffs,
           "max_abs_diff": max(diffs), "max_abs_diff_first10": max(diffs[:10]),
           "own_init_std_over_reference_init_std": init_stats,
           "config": {"layers": a.layers, "hidden": a.hidden, "heads": a.heads, "ffn": a.ffn, "seq": a.seq, "vocab": a.vocab,
                      "micro_batch": a.micro_batch, "n_mb": a.n_mb}}
    if rank == 0:
        os.makedirs(os.path.dirname(a.out), exist_ok=True)
        json.dump(res, open(a.out, "w"), indent=1)
        print("REF_PARITY " + json.dumps(res), flush=True)
    import torch.distributed as dist

    if dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
