"""GPU numerics: every sm_100a kernel vs a plain PyTorch fp32 reference of the same op."""
import importlib.util
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

_spec = importlib.util.spec_from_file_location(
    "gpu_checks", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "gpu_checks.py"))
gc = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(gc)


def _assert_small(res, tol):
    bad = {k: v for k, v in res.items() if not k.startswith("_") and (isinstance(v, str) or v > tol)}
    assert not bad, bad
    assert not res.get("_fallbacks"), f"library fallback used: {res['_fallbacks']}"


def test_extension_loaded():
    from megatronapp_b200 import ops

    assert ops.ext() is not None


@pytest.mark.parametrize("mode", ["nt", "nn", "tn", "tt", "tt_accum", "nt_bias_gelu"])
def test_gemm_layouts(mode):
    res = gc.check_gemm(mode, [(256, 256, 128), (384, 512, 320), (2048, 4096, 4096), (300, 520, 264)])
    _assert_small(res, 2e-2)


def test_gemm_fp16():
    _assert_small(gc.check_gemm("nt", [(256, 256, 128), (1024, 2048, 512)], "fp16"), 5e-3)


def test_elementwise_kernels():
    res = gc.check_elementwise()
    keep = res.pop("bda_keep_frac(expect .75)")
    assert abs(keep - 0.75) < 0.01
    _assert_small(res, 3e-2)


def test_attention_kernels():
    from megatronapp_b200 import ops

    if not ops.native_attention_available():
        pytest.skip("native attention kernel not built in this revision")
    res = gc.check_attention()
    res.pop("native_available")
    _assert_small(res, 3e-2)


def test_model_step_matches_fp32_reference():
    res = gc.check_model_parity()
    assert res["max_abs_diff"] < 0.08, res
    assert res["fallbacks"].get("gemm", 0) == 0, res


def test_smoke():
    gc.check_smoke()
