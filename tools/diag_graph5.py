import os, sys, subprocess, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

def run(case):
    import torch
    from megatronapp_b200.config import TransformerConfig
    from megatronapp_b200.models.gpt import get_gpt_layer_local_spec
    from megatronapp_b200.parallel import state as ps
    from megatronapp_b200.parallel.random import model_parallel_cuda_manual_seed
    from megatronapp_b200.transformer.cuda_graphs import CudaGraphManager, _TensorOnly
    from megatronapp_b200.transformer.spec_utils import build_module

    ps.initialize_model_parallel(1, 1)
    model_parallel_cuda_manual_seed(1)
    cfg = TransformerConfig(num_layers=1, hidden_size=512, num_attention_heads=4, bf16=True, params_dtype=torch.bfloat16,
                            hidden_dropout=0.0, attention_dropout=0.0)
    layer = build_module(get_gpt_layer_local_spec(), config=cfg, layer_number=1).cuda()
    mk = lambda: torch.randn(256, 2, 512, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    if case == "direct":            # make_graphed_callables on _TensorOnly directly, one eager step before
        x = mk()
        _TensorOnly(layer)(x).float().pow(2).sum().backward()
        g = torch.cuda.make_graphed_callables(_TensorOnly(layer), (torch.zeros_like(x).requires_grad_(True),), num_warmup_iters=3)
        g(mk()).float().pow(2).sum().backward()
    elif case == "mgr_only":        # manager, no eager comparison in between
        mgr = CudaGraphManager(layer, warmup_steps=1)
        for i in range(3):
            mgr(mk()).float().pow(2).sum().backward()
    elif case == "mgr_seed":        # manager + torch.manual_seed in the loop
        mgr = CudaGraphManager(layer, warmup_steps=1)
        for i in range(3):
            torch.manual_seed(i)
            mgr(mk()).float().pow(2).sum().backward()
    elif case == "mgr_eager_between":
        mgr = CudaGraphManager(layer, warmup_steps=1)
        for i in range(3):
            x = mk()
            mgr(x).float().pow(2).sum().backward()
            xe = x.detach().clone().requires_grad_(True)
            layer(xe)[0].float().pow(2).sum().backward()
    torch.cuda.synchronize()
    print(case, "OK", flush=True)

if __name__ == "__main__":
    if len(sys.argv) > 1:
        try:
            run(sys.argv[1])
        except Exception:
            traceback.print_exc()
            sys.exit(1)
    else:
        for c in ("direct", "mgr_only", "mgr_seed", "mgr_eager_between"):
            p = subprocess.run([sys.executable, os.path.abspath(__file__), c], capture_output=True, text=True, timeout=180)
            lines = [l for l in (p.stdout + p.stderr).splitlines() if l.strip()]
            keep = [l for l in lines if " OK" in l or "Error" in l or "megatronapp" in l][:6]
            print(f"[{'OK' if p.returncode == 0 else 'FAIL'}] {c}: " + " | ".join(keep)[:600], flush=True)
