import os, sys, subprocess, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

def run(case):
    import torch
    from megatronapp_b200 import ops

    class LN(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.g = torch.nn.Parameter(torch.ones(512, device="cuda", dtype=torch.bfloat16))
        def forward(self, x):
            return ops.layer_norm(x, self.g, None, 1e-5, rms=True)

    class Lin(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.randn(512, 512, device="cuda", dtype=torch.bfloat16) * 0.02)
        def forward(self, x):
            return ops._LinearFn.apply(x, self.w, None, False)

    class Scale(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.ones(512, device="cuda", dtype=torch.bfloat16))
        def forward(self, x):
            return x * self.w

    m = {"torch_linear": lambda: torch.nn.Linear(512, 512, device="cuda", dtype=torch.bfloat16), "scale": Scale, "ln": LN, "lin": Lin}[case]()
    x = torch.randn(256, 512, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    g = torch.cuda.make_graphed_callables(m, (x,), num_warmup_iters=3)
    for i in range(3):
        xi = torch.randn(256, 512, device="cuda", dtype=torch.bfloat16, requires_grad=True)
        y = g(xi)
        y.float().sum().backward()
    torch.cuda.synchronize()
    print(case, "make_graphed_callables OK", float(xi.grad.float().abs().sum()), flush=True)

if __name__ == "__main__":
    if len(sys.argv) > 1:
        try:
            run(sys.argv[1])
        except Exception:
            traceback.print_exc()
            sys.exit(1)
    else:
        for c in ("torch_linear", "scale", "ln", "lin"):
            p = subprocess.run([sys.executable, os.path.abspath(__file__), c], capture_output=True, text=True, timeout=180)
            lines = [l for l in (p.stdout + p.stderr).splitlines() if l.strip()]
            keep = [l for l in lines if "OK" in l or "Error" in l or "File \"/root" in l or "megatronapp" in l][:10]
            print(f"[{'OK' if p.returncode == 0 else 'FAIL'}] {c}: " + " | ".join(keep)[:900], flush=True)
