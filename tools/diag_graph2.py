import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

def run(case, mode):
    import torch
    from megatronapp_b200 import ops

    class PureFn(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x):
            ctx.save_for_backward(x)
            return x * 2
        @staticmethod
        def backward(ctx, g):
            (x,) = ctx.saved_tensors
            print("  bwd thread current stream:", torch.cuda.current_stream(), "capturing:", torch.cuda.is_current_stream_capturing(), flush=True)
            return g * 2

    class ExtFn(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x):
            return x.clone()
        @staticmethod
        def backward(ctx, g):
            print("  bwd thread current stream:", torch.cuda.current_stream(), "capturing:", torch.cuda.is_current_stream_capturing(), flush=True)
            gg = g.contiguous().clone()
            ops.ext().scale_(gg, 2.0, None)
            return gg

    x = torch.randn(256, 512, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    g = torch.ones(512, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    f = {"pure": lambda: PureFn.apply(x), "ext": lambda: ExtFn.apply(x), "ln": lambda: ops.layer_norm(x, g, None, 1e-5, rms=True)}[case]
    for _ in range(2):
        f().float().sum().backward()
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    print("capture stream:", s, flush=True)
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        with torch.cuda.graph(gr, stream=s, capture_error_mode=mode):
            y = f()
            grads = torch.autograd.grad(y.float().sum(), [x], allow_unused=True)
    gr.replay()
    torch.cuda.synchronize()
    print(case, mode, "ALL OK", flush=True)

if __name__ == "__main__":
    if len(sys.argv) > 2:
        run(sys.argv[1], sys.argv[2])
    else:
        for c in ("pure", "ext", "ln"):
            for m in ("global", "thread_local", "relaxed"):
                p = subprocess.run([sys.executable, os.path.abspath(__file__), c, m], capture_output=True, text=True, timeout=180)
                lines = [l for l in (p.stdout + p.stderr).splitlines() if l.strip()]
                keep = [l for l in lines if "stream" in l or "OK" in l or "Error" in l][:5]
                print(f"[{'OK' if p.returncode == 0 else 'FAIL'}] {c}/{m}: " + " | ".join(keep)[:600], flush=True)
