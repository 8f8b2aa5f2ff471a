"""2-GPU numerics of the fused compute+collective kernels (skipped on boxes with one GPU)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_fused_collective_kernels_two_gpus():
    out = os.path.join(ROOT, "gpurun_out", "multi_pytest.json")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29547", os.path.join(ROOT, "tools", "multi_gpu_checks.py"), "--out", out, "--skip-model"]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert p.returncode == 0, (p.stdout[-2000:], p.stderr[-3000:])
    res = json.load(open(out))
    assert not [k for k in res if k.startswith("EXCEPTION")], {k: v for k, v in res.items() if k.startswith("EXCEPTION")}
    errs = {k: v for k, v in res.items() if k.endswith("_err") and isinstance(v, float)}
    assert errs and max(errs.values()) < 3e-2, {k: v for k, v in errs.items() if v >= 3e-2}
    assert all(v for k, v in res.items() if k.endswith("_deterministic"))
    moe = res.get("moe_ep_err")
    assert moe is not None and not moe["overflow"] and max(v for k, v in moe.items() if k != "overflow") < 6e-2, moe
