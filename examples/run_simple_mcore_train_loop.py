"""Minimal training loop on the core API only (no CLI, no global args): the reference's
``examples/run_simple_mcore_train_loop.py`` is broken as shipped because its core imports the
training harness; here the same flow works on CPU/gloo and on GPUs.

    torchrun --nproc-per-node 1 --master-addr 127.0.0.1 examples/run_simple_mcore_train_loop.py [--trace-dir DIR]
"""
import argparse
import os
import sys
from functools import partial

if not os.path.exists("/dev/nvidia0"):   # CPU/gloo mode: avoid OpenMP oversubscription
    os.environ.setdefault("OMP_NUM_THREADS", "2")

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from megatronapp_b200.config import TransformerConfig  # noqa: E402
from megatronapp_b200.models.gpt import GPTModel, get_gpt_layer_local_spec  # noqa: E402
from megatronapp_b200.parallel import state as parallel_state  # noqa: E402
from megatronapp_b200.parallel.random import model_parallel_cuda_manual_seed  # noqa: E402
from megatronapp_b200.parallel.schedules import get_forward_backward_func  # noqa: E402
from megatronapp_b200.scan.tracer import get_tracer  # noqa: E402


def initialize_distributed(tp=1, pp=1):
    parallel_state.init_distributed()
    parallel_state.initialize_model_parallel(tp, pp)


def model_provider(device):
    cfg = TransformerConfig(num_layers=12, hidden_size=768, num_attention_heads=12, use_cpu_initialization=device == "cpu",
                            pipeline_dtype=torch.float32, hidden_dropout=0.0, attention_dropout=0.0)
    return GPTModel(cfg, get_gpt_layer_local_spec(), vocab_size=50304, max_sequence_length=64).to(device)


def forward_step_func(data_iterator, model):
    def loss_func(loss_mask, output_tensor):
        losses = output_tensor.float().view(-1)
        loss = torch.sum(losses * loss_mask.view(-1)) / loss_mask.sum()
        return loss, {"lm loss": loss.detach()}

    d = next(data_iterator)
    dev = next(model.parameters()).device
    out = model(d["tokens"].to(dev), d["position_ids"].to(dev), None, labels=d["labels"].to(dev))
    return out, partial(loss_func, d["loss_mask"].to(dev))


def batches(seq=64, bsz=2, vocab=50304):
    g = torch.Generator().manual_seed(0)
    while True:
        t = torch.randint(0, vocab, (bsz, seq + 1), generator=g)
        yield {"tokens": t[:, :-1], "labels": t[:, 1:], "loss_mask": torch.ones(bsz, seq),
               "position_ids": torch.arange(seq).unsqueeze(0).expand(bsz, -1)}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--trace-dir", default=None)
    a = ap.parse_args()
    initialize_distributed()
    model_parallel_cuda_manual_seed(123)
    device = "cuda" if torch.cuda.is_available() else "cpu"
    model = model_provider(device)      # GPT-2 125M, TP1 PP1 DP1
    optim = torch.optim.Adam(model.parameters(), lr=1e-4)
    tracer = get_tracer()
    if a.trace_dir:
        tracer.configure(True, a.trace_dir, interval=1, continuous_iterations=1)
    fb = get_forward_backward_func()
    it = batches()
    for i in range(a.iters):
        tracer.iteration_begin(i)
        optim.zero_grad(set_to_none=False)
        losses = fb(forward_step_func=forward_step_func, data_iterator=it, model=model, num_microbatches=1,
                    seq_length=64, micro_batch_size=2, forward_only=False)
        with tracer.scope("optimizer"):
            optim.step()
        tracer.iteration_end()
        print(f"iteration {i}: loss {float(losses[0]['lm loss']):.4f}", flush=True)
    if a.trace_dir:
        tracer.shutdown()
        from megatronapp_b200.scan.aggregate import aggregate

        traces, _ = aggregate(a.trace_dir, os.path.join(a.trace_dir, "chrome_trace.json"))
        print(f"MegaScan: wrote {len(traces)} events to {a.trace_dir}/chrome_trace.json")
