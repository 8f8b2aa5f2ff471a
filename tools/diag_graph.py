"""Which op breaks CUDA-graph capture?  Captures forward+backward of each op in its own subprocess."""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

CASES = ["layer_norm", "rms_norm", "linear", "linear_fused_wgrad", "bias_gelu", "fused_gelu_linear", "bias_dropout_add",
         "attention", "swiglu", "rope", "cross_entropy", "layer"]


def run(case):
    import torch

    from megatronapp_b200 import ops

    dev = "cuda"
    from megatronapp_b200.parallel import state as _ps
    from megatronapp_b200.parallel.random import model_parallel_cuda_manual_seed as _seed

    _ps.initialize_model_parallel(1, 1)
    _seed(1)
    torch.manual_seed(0)
    x = torch.randn(256, 2, 512, device=dev, dtype=torch.bfloat16, requires_grad=True)
    w = torch.randn(1024, 512, device=dev, dtype=torch.bfloat16, requires_grad=True)
    b = torch.randn(1024, device=dev, dtype=torch.bfloat16, requires_grad=True)
    g = torch.ones(512, device=dev, dtype=torch.bfloat16, requires_grad=True)
    be = torch.zeros(512, device=dev, dtype=torch.bfloat16, requires_grad=True)

    def f():
        if case == "layer_norm":
            return ops.layer_norm(x, g, be, 1e-5)
        if case == "rms_norm":
            return ops.layer_norm(x, g, None, 1e-5, rms=True)
        if case == "linear":
            return ops._LinearFn.apply(x, w, b, False) if hasattr(ops, "_LinearFn") else None
        if case == "linear_fused_wgrad":
            return ops._LinearFn.apply(x, w, b, True)
        if case == "bias_gelu":
            return ops.bias_gelu(torch.nn.functional.linear(x, w), b)
        if case == "fused_gelu_linear":
            from megatronapp_b200.parallel.tp_layers import _ColumnParallelFn

            return _ColumnParallelFn.apply(x, w, b, False, False, False, 1)
        if case == "bias_dropout_add":
            return ops.bias_dropout_add(x, be, x, 0.0, True)
        if case == "attention":
            q = x.view(256, 2, 4, 128)
            return ops.attention(q.permute(1, 2, 0, 3), q.permute(1, 2, 0, 3), q.permute(1, 2, 0, 3), None, True, 0.0, True, None)
        if case == "swiglu":
            return ops.swiglu(torch.nn.functional.linear(x, w), None)
        if case == "rope":
            cs = ops.rope_cos_sin(256, 128, device=dev)
            return ops.rope(x.view(256, 2, 4, 128), cs, False)
        if case == "cross_entropy":
            lab = torch.randint(0, 1024, (256, 2), device=dev)
            return ops.cross_entropy(torch.nn.functional.linear(x, w), lab)
        if case == "layer":
            from megatronapp_b200.config import TransformerConfig
            from megatronapp_b200.models.gpt import get_gpt_layer_local_spec
            from megatronapp_b200.parallel import state as ps
            from megatronapp_b200.parallel.random import model_parallel_cuda_manual_seed
            from megatronapp_b200.transformer.spec_utils import build_module

            if not hasattr(run, "layer"):
                cfg = TransformerConfig(num_layers=1, hidden_size=512, num_attention_heads=4, bf16=True, params_dtype=torch.bfloat16,
                                        hidden_dropout=0.0, attention_dropout=0.0)
                run.layer = build_module(get_gpt_layer_local_spec(), config=cfg, layer_number=1).cuda()
            return run.layer(x)[0]

    for _ in range(2):       # warm-up (probes, attribute setting, lazy inits) outside capture
        y = f()
        y.float().sum().backward()
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        with torch.cuda.graph(gr, stream=s):
            y = f()
        print(case, "forward capture OK", flush=True)
        gb = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gb, stream=s, pool=gr.pool()):
            grads = torch.autograd.grad(y.float().sum(), [t for t in (x, w, b, g, be) if t.requires_grad], allow_unused=True)
        print(case, "backward capture OK", flush=True)
    gr.replay()
    gb.replay()
    torch.cuda.synchronize()
    print(case, "replay OK", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run(sys.argv[1])
    else:
        for c in CASES:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), c], capture_output=True, text=True, timeout=180,
                               env=dict(os.environ))
            lines = [l for l in (p.stdout + p.stderr).splitlines() if l.strip()]
            keep = [l for l in lines if " OK" in l or "Error" in l or "error" in l][:6]
            status = "OK" if p.returncode == 0 else "FAIL"
            print(f"[{status}] {c}: " + " | ".join(keep)[:700], flush=True)
