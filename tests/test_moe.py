import pytest
import torch

from dist_utils import run_distributed


def _close(a, b, tol=1e-5):
    assert max(abs(x - y) for x, y in zip(a, b)) < tol


@pytest.mark.parametrize("dispatcher", ["alltoall", "allgather"])
def test_expert_parallel_matches_single_rank(dispatcher):
    # EP=2: each rank routes its own tokens; compare against EP=1 runs fed the same per-rank tokens
    ep2 = run_distributed("moe_parity", 2, ep=2, dispatcher=dispatcher)
    ref = run_distributed("moe_parity", 2, ep=1, dispatcher="alltoall")   # dp=2: same seeds per rank
    for r in range(2):
        _close(ep2[r]["out"], ref[r]["out"])
        _close(ep2[r]["gx"], ref[r]["gx"])


def test_grouped_mlp_matches_sequential():
    a = run_distributed("moe_parity", 1, ep=1, grouped=False)[0]
    b = run_distributed("moe_parity", 1, ep=1, grouped=True)[0]
    # same weights by construction only if layouts agree: fc1 [gate|up] rows, fc2 [h, ffn]
    _close(a["out"], b["out"])


def test_flex_dispatcher_single_rank_matches_alltoall():
    a = run_distributed("moe_parity", 1, ep=1, dispatcher="alltoall")[0]
    b = run_distributed("moe_parity", 1, ep=1, dispatcher="flex")[0]
    _close(a["out"], b["out"])
    _close(a["gx"], b["gx"])


@pytest.mark.parametrize("grouped", [True, False])
def test_fused_dispatcher_matches_alltoall(grouped):
    """Ragged 256-row padded layout + grouped GEMM (reference implementation of the CUDA path) == per-expert modules."""
    a = run_distributed("moe_parity", 1, ep=1, dispatcher="alltoall")[0]
    b = run_distributed("moe_parity", 1, ep=1, dispatcher="fused", grouped=grouped)[0]
    _close(a["out"], b["out"])
    _close(a["gx"], b["gx"])
    _close(a["grouter"], b["grouter"])


def test_expert_parallel_grad_norm_and_clipping_agree_across_ranks():
    res = run_distributed("ep_clip_consistency", 2)
    assert res[0]["norms"] and res[0]["norms"] == pytest.approx(res[1]["norms"], rel=1e-6), res
    assert res[0]["dense"] == pytest.approx(res[1]["dense"], rel=1e-9), res      # dense replicas stay identical
