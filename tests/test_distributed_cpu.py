"""Multi-process (gloo, world_size > 1) checks of the parallel runtime on CPU."""
import math

import pytest

from dist_utils import run_distributed


def _ref_losses():
    return run_distributed("train_parity", 1)[0]["losses"]


@pytest.fixture(scope="module")
def ref():
    return _ref_losses()


def _last_stage_losses(res):
    """Mean over the data-parallel replicas' last-stage losses (each replica reports its own shard)."""
    rows = [r["losses"] for r in res
            if r["last"] and r.get("is_backward", True) and not any(math.isnan(x) for x in r["losses"])]
    if not rows:
        raise AssertionError(res)
    return [sum(col) / len(col) for col in zip(*rows)]


def _close(a, b, tol):
    assert len(a) == len(b)
    for x, y in zip(a, b):
        assert abs(x - y) < tol, (a, b)


def test_collectives(ref):
    for r in run_distributed("collectives_roundtrip", 2):
        assert r["gather_shape"] == [8, 2] and r["rs_ok"] and r["a2a_roundtrip"]


def test_vocab_parallel_cross_entropy():
    for r in run_distributed("vocab_parallel_ce", 2):
        assert r["fwd"] < 1e-5 and r["bwd"] < 1e-6


def test_vocab_parallel_distillation_kl():
    for r in run_distributed("vocab_parallel_kd_loss", 2):
        assert max(r["fwd0"], r["fwd1"]) < 1e-5 and max(r["bwd0"], r["bwd1"]) < 1e-6, r


def test_data_parallel_matches_single(ref):
    res = run_distributed("train_parity", 2)
    _close(_last_stage_losses(res), ref, 2e-4)
    assert abs(res[0]["w"] - res[1]["w"]) < 1e-9


def test_zero1_distributed_optimizer(ref):
    res = run_distributed("train_parity", 2, dist_opt=True)
    _close(_last_stage_losses(res), ref, 2e-4)
    assert abs(res[0]["w"] - res[1]["w"]) < 1e-9


def test_overlap_grad_reduce(ref):
    res = run_distributed("train_parity", 2, overlap=True)
    _close(_last_stage_losses(res), ref, 2e-4)


def test_pipeline_1f1b(ref):
    # different init (per-stage seeds) => compare to itself across schedules instead of to TP1
    a = _last_stage_losses(run_distributed("train_parity", 2, pp=2))
    assert all(math.isfinite(x) for x in a) and a[-1] < a[0]


def test_pp2_dp2_distributed_optimizer_keeps_tied_embeddings_in_sync():
    """PP2 x DP2 + ZeRO-1 with tied embeddings: the embedding weight sits ALONE in its bucket on the first and the last stage
    (identical shard ranges), so the first<->last embedding all-reduce on reduce-scattered shards is exact: same losses as
    PP2 x DP2 with the replicated optimizer, and the two copies of the embedding stay bit-identical."""
    base = run_distributed("train_parity", 4, pp=2)
    zero = run_distributed("train_parity", 4, pp=2, dist_opt=True)
    _close(_last_stage_losses(zero), _last_stage_losses(base), 1e-5)
    embs = [r["emb"] for r in zero if r.get("emb") is not None]
    assert len(embs) == 4 and max(embs) - min(embs) < 1e-9, embs


def test_pipeline_interleaved_and_dpp_agree():
    base = _last_stage_losses(run_distributed("train_parity", 2, pp=2, vpp=2))
    for pol in ("static", "bfc", "dfc", "auto"):
        got = _last_stage_losses(run_distributed("train_parity", 2, pp=2, vpp=2, dpp=pol))
        _close(got, base, 1e-5)


def test_tensor_parallel_sequence_parallel_trains():
    a = _last_stage_losses(run_distributed("train_parity", 2, tp=2, sp=True))
    b = _last_stage_losses(run_distributed("train_parity", 2, tp=2, sp=False))
    _close(a, b, 2e-4)   # SP and non-SP are the same math
    assert a[-1] < a[0]


def test_fbd_matches_plain_pipeline():
    """4 ranks: pp=2 x (forward, backward) twins == 2 ranks pp=2 (same data, same seeds)."""
    base = _last_stage_losses(run_distributed("train_parity", 2, pp=2))
    res = run_distributed("train_parity", 4, pp=2, fbd=True)
    _close(_last_stage_losses(res), base, 1e-5)


def test_fbd_interleaved_matches_plain_interleaved_pipeline():
    """MegaFBD on the virtual-pipeline layout: 4 ranks = pp 2 x (forward, backward) twins with 2 chunks per rank reproduce
    the 2-rank interleaved pipeline (same data, same seeds) -- activations of a group of pp microbatches are handed over in
    forward order and consumed by the backward twin walking the virtual stages in reverse."""
    base = _last_stage_losses(run_distributed("train_parity", 2, pp=2, vpp=2))
    res = run_distributed("train_parity", 4, pp=2, vpp=2, fbd=True)
    _close(_last_stage_losses(res), base, 1e-5)
    assert sum(1 for r in res if not r["is_backward"]) == 2


@pytest.mark.parametrize("swiglu", [False, True])
def test_fbd_forward_tp_collapse_matches_plain_tp_pipeline(swiglu):
    """--ignore-forward-tensor-parallel: 6 ranks = pp 2 x (1 un-sharded forward rank + 2 backward TP ranks, sequence
    parallel) reproduce the losses of the 4-rank tp2 x pp2 pipeline: the forward rank runs true TP=1 math on weights
    re-assembled from its twins' shards every step (gated fc1 included) and hands each twin its s/tp slice."""
    base = _last_stage_losses(run_distributed("train_parity", 4, tp=2, pp=2, sp=True, swiglu=swiglu))
    res = run_distributed("train_parity", 6, tp=2, pp=2, sp=True, fbd=True, collapse=True, swiglu=swiglu)
    _close(_last_stage_losses(res), base, 1e-4)
    assert sum(1 for r in res if not r["is_backward"]) == 2


@pytest.mark.parametrize("comm", ["a2a", "all_gather", "p2p"])
def test_context_parallel_attention(comm):
    for r in run_distributed("cp_attention_parity", 2, comm=comm):
        assert max(r["fwd"], r["dq"], r["dk"], r["dv"]) < 1e-5, r   # reference attention runs in fp32


@pytest.mark.parametrize("causal", [True, False])
def test_hierarchical_context_parallel_attention(causal):
    """cp 4 = 2 (all-to-all inside) x 2 (K/V ring outside): forward and all three gradients match single-rank attention"""
    for r in run_distributed("cp_attention_parity", 4, comm="a2a+p2p", hcp=[2, 2], causal=causal):
        assert max(r["fwd"], r["dq"], r["dk"], r["dv"]) < 1e-5, r


def test_fsdp_zero3_matches_unsharded_training():
    for r in run_distributed("fsdp_parity", 2):
        _close(r["losses"], r["ref"], 2e-4)
        assert r["param_err"] < 2e-4
        assert r["resident_bytes_after_step"] == 0            # every unit is freed between steps
        assert 0.5 <= r["shard_fraction"] < 0.56               # each rank owns ~1/2 of the parameters


@pytest.mark.parametrize("comm", ["a2a", "p2p"])
def test_context_parallel_gpt_matches_single_rank(comm):
    for r in run_distributed("cp_model_parity", 2, comm=comm):
        assert abs(r["loss_full"] - r["loss_cp"]) < 1e-4, r
        assert r["grad_err"] < 1e-4 * max(1.0, r["grad_ref_max"]) + 1e-5, r


@pytest.mark.parametrize("sp", [False, True])
def test_mamba_mixer_tensor_parallel_and_reshard(tmp_path, sp):
    res = run_distributed("mamba_tp_reshard", 2, path=str(tmp_path / "ck"), sp=sp)
    for r in res:
        assert r["local_heads"] * 2 == r["full_heads"]
        assert r["fwd"] < 1e-4 * max(1.0, r["scale"]) and r["gx"] < 1e-3, r


def test_t5_encoder_decoder_pipeline_split_matches_unsplit():
    base = run_distributed("t5_pp_parity", 1)[0]
    assert len(base["losses"]) == 2 and base["losses"][1] < base["losses"][0]
    for pp, split in ((2, 1), (4, 2), (3, 1)):
        res = run_distributed("t5_pp_parity", pp, pp=pp, split=split)
        last = [r for r in res if r["losses"]][0]
        assert last["losses"] == pytest.approx(base["losses"], rel=1e-5), (pp, split, last["losses"], base["losses"])
        # the tied word embedding (stage 0 / first decoder stage) and the shared position table stay identical copies
        embs = [r["emb"] for r in res if r["emb"] is not None]
        assert len(embs) == 2 and embs[0] == pytest.approx(embs[1], rel=1e-9) and embs[0] == pytest.approx(base["emb"], rel=1e-6)
        poss = [r["pos"] for r in res if r["pos"] is not None]
        assert poss[0] == pytest.approx(poss[1], rel=1e-9) and poss[0] == pytest.approx(base["pos"], rel=1e-6)


def test_generation_under_tensor_and_pipeline_parallelism():
    base = run_distributed("generation_parity", 1)[0]
    assert all(1 <= len(t) <= 6 for t in base["tokens"]) and max(len(t) for t in base["tokens"]) > 2, base["tokens"]
    assert base["dynamic"] == base["tokens"]
    for kw in ({"tp": 2}, {"pp": 2}):
        res = run_distributed("generation_parity", 2, **kw)
        for r in res:
            assert r["tokens"] == base["tokens"], (kw, r["rank"])          # every rank ends with the same continuation
            assert r["dynamic"] == base["tokens"], (kw, r["rank"])         # paged-KV continuous batching too
        lp = [r["logprobs"] for r in res if r["logprobs"][0]]
        assert lp and all(abs(a - b) < 1e-3 for x, y in zip(lp[-1], base["logprobs"]) for a, b in zip(x, y))


def test_sharded_checkpoint_reshards_across_tp_and_pp(tmp_path):
    """A checkpoint written at tp=2 loads at tp=1, at pp=2 (tied embedding copy on the last stage, global layer ids) and
    back at tp=2 -- every layout reproduces the saver's logits."""
    import numpy as np

    path = str(tmp_path / "ck")
    saved = run_distributed("reshard_logits", 2, tp=2, path=path, mode="save")
    ref = np.asarray(saved[0]["logits"])
    assert np.allclose(ref, np.asarray(saved[1]["logits"]))
    for world, kw in ((1, {}), (2, {"pp": 2}), (2, {"tp": 2})):
        res = run_distributed("reshard_logits", world, path=path, mode="load", **kw)
        got = [np.asarray(r["logits"]) for r in res if r["logits"] is not None]
        assert got and all(np.allclose(g, ref, atol=1e-5) for g in got), kw
    # and the other direction: written by a 2-stage pipeline, read by a single rank
    path2 = str(tmp_path / "ck_pp")
    saved = run_distributed("reshard_logits", 2, pp=2, path=path2, mode="save")
    ref2 = np.asarray([r["logits"] for r in saved if r["logits"] is not None][0])
    one = run_distributed("reshard_logits", 1, path=path2, mode="load")[0]
    assert np.allclose(np.asarray(one["logits"]), ref2, atol=1e-5)


@pytest.mark.parametrize("swiglu", [False, True])
def test_torch_dist_checkpoint_reshards_across_tp_and_pp(tmp_path, swiglu):
    """Reference-layout (DCP) checkpoints: stacked layer keys, gate/up-split gated fc1 -- written at tp=2, read at tp=1 and
    pp=2 and the other way round; every layout reproduces the saver's logits (the swiglu case is the one a rank-major
    concatenation of [gate_r; up_r] shards gets wrong)."""
    import numpy as np

    path = str(tmp_path / "ck")
    saved = run_distributed("reshard_logits", 2, tp=2, path=path, mode="save", fmt="torch_dist", swiglu=swiglu)
    ref = np.asarray(saved[0]["logits"])
    for world, kw in ((1, {}), (2, {"pp": 2})):
        res = run_distributed("reshard_logits", world, path=path, mode="load", fmt="torch_dist", swiglu=swiglu, **kw)
        got = [np.asarray(r["logits"]) for r in res if r["logits"] is not None]
        assert got and all(np.allclose(g, ref, atol=1e-5) for g in got), kw
    path1 = str(tmp_path / "ck1")
    one = run_distributed("reshard_logits", 1, path=path1, mode="save", fmt="torch_dist", swiglu=swiglu)[0]
    two = run_distributed("reshard_logits", 2, tp=2, path=path1, mode="load", fmt="torch_dist", swiglu=swiglu)
    assert np.allclose(np.asarray(two[0]["logits"]), np.asarray(one["logits"]), atol=1e-5)


def test_moe_checkpoint_reshards_over_expert_parallel_and_grouped(tmp_path):
    import numpy as np

    path = str(tmp_path / "moe")
    saved = run_distributed("moe_reshard", 2, ep=2, path=path, mode="save")
    assert saved[0]["keys"][0].startswith("mlp.experts.local_experts.0.") and saved[1]["keys"][0].startswith("mlp.experts.local_experts.2.")
    ref = np.asarray(saved[0]["out"])
    assert np.allclose(ref, np.asarray(saved[1]["out"]), atol=1e-6)
    for world, kw in ((1, dict(ep=1)), (1, dict(ep=1, grouped=True)), (2, dict(ep=2, grouped=True))):
        res = run_distributed("moe_reshard", world, path=path, mode="load", **kw)
        for r in res:
            assert np.allclose(np.asarray(r["out"]), ref, atol=1e-5), kw


def test_local_checkpoint_manager_replication_survives_a_lost_node(tmp_path):
    res = run_distributed("local_ckpt_replication", 3, path=str(tmp_path))
    for r, o in enumerate(res):
        assert o["holders"] == [(r + 1) % 3] and o["sources"] == [(r - 1) % 3]
        assert o["kept"] == ["iter_0000004"]                                   # older iterations are cleaned up
        assert o["latest"] == 4 and o["latest_after_loss"] == 4
        assert o["restored"] == [4, float(10 * r + 4)] and o["own_file_back"]
        assert o["latest_after_two_losses"] == -1                              # rank 1's state had one holder (rank 2): both gone


def test_evidence_indexer_shards_over_ranks_and_feeds_the_mips_index(tmp_path):
    res = run_distributed("evidence_indexer", 2, path=str(tmp_path))
    r0 = [r for r in res if r["rank"] == 0][0]
    assert r0["ids"] == list(range(10, 17)) and r0["n"] == 7 and r0["tmp_gone"]
    assert r0["self_hits"] == r0["ids"]                     # every passage retrieves itself first
