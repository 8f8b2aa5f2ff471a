import os, sys, subprocess, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

def run(case, eager_first):
    import torch
    from megatronapp_b200 import ops
    from megatronapp_b200.config import TransformerConfig
    from megatronapp_b200.models.gpt import get_gpt_layer_local_spec
    from megatronapp_b200.parallel import state as ps
    from megatronapp_b200.parallel.random import model_parallel_cuda_manual_seed
    from megatronapp_b200.transformer.spec_utils import build_module

    ps.initialize_model_parallel(1, 1)
    model_parallel_cuda_manual_seed(1)
    cfg = TransformerConfig(num_layers=1, hidden_size=512, num_attention_heads=4, bf16=True, params_dtype=torch.bfloat16,
                            hidden_dropout=0.0, attention_dropout=0.0)
    layer = build_module(get_gpt_layer_local_spec(), config=cfg, layer_number=1).cuda()

    class Wrap(torch.nn.Module):
        def __init__(self, f, mods):
            super().__init__()
            self.f = f
            self.mods = torch.nn.ModuleList(mods)
        def forward(self, x):
            return self.f(x)

    def attn(x):
        o, b = layer.self_attention(x)
        return o + b
    def mlp(x):
        o, b = layer.mlp(x)
        return o + b
    def core(x):
        q = x.view(256, 2, 4, 128).permute(1, 2, 0, 3)
        return ops.attention(q, q, q, None, True, 0.0, True, None).permute(2, 0, 1, 3).reshape(256, 2, 512)
    def bda(x):
        return ops.bias_dropout_add(x, layer.mlp.linear_fc2.bias, x, 0.0, True)
    def full(x):
        return layer(x)[0]
    f, mods = {"attn": (attn, [layer.self_attention]), "mlp": (mlp, [layer.mlp]), "core": (core, []), "bda": (bda, [layer.mlp.linear_fc2]),
               "full": (full, [layer])}[case]
    m = Wrap(f, mods)
    x = torch.randn(256, 2, 512, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    if eager_first == "1":
        m(x).float().sum().backward()
    g = torch.cuda.make_graphed_callables(m, (torch.zeros_like(x).requires_grad_(True),), num_warmup_iters=3)
    for i in range(2):
        xi = torch.randn(256, 2, 512, device="cuda", dtype=torch.bfloat16, requires_grad=True)
        g(xi).float().sum().backward()
    torch.cuda.synchronize()
    print(case, eager_first, "OK", flush=True)

if __name__ == "__main__":
    if len(sys.argv) > 2:
        try:
            run(sys.argv[1], sys.argv[2])
        except Exception:
            traceback.print_exc()
            sys.exit(1)
    else:
        for c in ("core", "bda", "mlp", "attn", "full"):
            for e in ("0", "1"):
                p = subprocess.run([sys.executable, os.path.abspath(__file__), c, e], capture_output=True, text=True, timeout=180)
                lines = [l for l in (p.stdout + p.stderr).splitlines() if l.strip()]
                keep = [l for l in lines if " OK" in l or "Error" in l or "megatronapp" in l][:8]
                print(f"[{'OK' if p.returncode == 0 else 'FAIL'}] {c} eager_first={e}: " + " | ".join(keep)[:700], flush=True)
