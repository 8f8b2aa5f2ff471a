"""``torch_dist`` checkpoint format: PyTorch-DCP on-disk layout, interchange with the UNMODIFIED reference
(baseline/_ref ``megatron.core.dist_checkpointing``), resharding of model and optimizer state, async writer process."""
import os
import pickle
import subprocess
import sys
import textwrap

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")

from megatronapp_b200.ckpt import torch_dist as td  # noqa: E402
from megatronapp_b200.ckpt.sharded import ShardedObject, ShardedTensor  # noqa: E402


def _st(key, data, offsets=(), pre=(), flat=None, replica=(0, 0, 0)):
    from megatronapp_b200.ckpt.sharded import _from_offsets_with_prepend

    st = _from_offsets_with_prepend(key, data, pre, list(offsets), replica)
    if flat is not None:
        st.flattened_range = slice(*flat)
        st.data = data.reshape(-1)[flat[0]:flat[1]].clone()
    return st


def test_layout_is_plain_pytorch_dcp(tmp_path):
    """.metadata is a pickled DCP Metadata, tensors are torch.save records: torch.distributed.checkpoint reads it back."""
    import torch.distributed.checkpoint as dcp

    w = torch.arange(48, dtype=torch.float32).view(8, 6)
    ssd = {"w#0": _st("blk.weight", w[:4], offsets=[(0, 0, 2)]), "w#1": _st("blk.weight", w[4:], offsets=[(0, 1, 2)]),
           "l0": _st("decoder.layers.ln.weight", torch.ones(6), pre=[(0, 0, 2)]),
           "l1": _st("decoder.layers.ln.weight", torch.full((6,), 2.0), pre=[(0, 1, 2)]),
           "obj": ShardedObject("blk._extra_state", {"a": 1})}
    td.save(ssd, str(tmp_path), {"iteration": 3}).finalize()
    assert sorted(os.listdir(tmp_path)) == [".metadata", "__0_0.distcp", "common.pt", "metadata.json"]
    md = pickle.load(open(tmp_path / ".metadata", "rb"))
    assert tuple(md.state_dict_metadata["blk.weight"].size) == (8, 6)
    assert tuple(md.state_dict_metadata["decoder.layers.ln.weight"].size) == (2, 6)
    assert "blk._extra_state/shard_0_1" in md.state_dict_metadata
    out = {"blk.weight": torch.zeros(8, 6), "decoder.layers.ln.weight": torch.zeros(2, 6)}
    dcp.load(out, checkpoint_id=str(tmp_path), no_dist=True)
    assert torch.equal(out["blk.weight"], w) and out["decoder.layers.ln.weight"][1].eq(2).all()


def test_reshard_on_load_including_flattened_optimizer_ranges(tmp_path):
    w = torch.randn(8, 6)
    m = torch.randn(8, 6)
    # written as 2 TP shards along axis 0; the optimizer moment as 2 x 2 flattened pieces (ZeRO-1 over dp = 2)
    ssd = {}
    for r in range(2):
        ssd[f"w{r}"] = _st("a.weight", w[4 * r:4 * r + 4], offsets=[(0, r, 2)])
        for d, (lo, hi) in enumerate(((0, 10), (10, 24))):
            ssd[f"m{r}{d}"] = _st("optimizer.state.exp_avg.a.weight", m[4 * r:4 * r + 4], offsets=[(0, r, 2)], flat=(lo, hi))
    td.save(ssd, str(tmp_path)).finalize()
    # read back under tp = 1, dp = 3 (three uneven flattened runs of the whole tensor) and tp = 4
    got_w = _st("a.weight", torch.zeros(8, 6))
    pieces = [_st("optimizer.state.exp_avg.a.weight", torch.zeros(8, 6), flat=f) for f in ((0, 7), (7, 30), (30, 48))]
    q = [_st("a.weight", torch.zeros(2, 6), offsets=[(0, r, 4)]) for r in range(4)]
    td.load({"w": got_w, **{f"p{i}": p for i, p in enumerate(pieces)}, **{f"q{i}": t for i, t in enumerate(q)}}, str(tmp_path))
    assert torch.equal(got_w.data, w)
    assert torch.equal(torch.cat([p.data for p in pieces]), m.reshape(-1))
    assert torch.equal(torch.cat([t.data for t in q]), w)


_REF_SNIPPET = textwrap.dedent('''
    import os, sys, types
    sys.path.insert(0, {ref!r})
    for n in ("shm_tensor_new_rdma", "shm_tensor_new_rdma_pre_alloc"):
        sys.modules.setdefault(n, types.ModuleType(n))
    import torch, torch.distributed as dist
    if not torch.cuda.is_available():
        torch.cuda.synchronize = lambda *a, **k: None    # the reference's writer syncs the device unconditionally
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", "{port}"
    dist.init_process_group("gloo", rank=0, world_size=1)
    from megatron.core import dist_checkpointing as dc
    from megatron.core.dist_checkpointing import ShardedTensor, ShardedObject
    torch.manual_seed(0)
    w = torch.arange(96, dtype=torch.float32).view(2, 8, 6)        # [layers, out, in]
    emb = torch.arange(40, dtype=torch.float32).view(10, 4)
    mode, path = sys.argv[1], sys.argv[2]
    def build(fill):
        sd = {{}}
        for l in range(2):
            for r in range(2):     # two TP shards per layer, layers stacked under one key (prepend axis)
                t = w[l, 4 * r:4 * r + 4].clone() if fill else torch.zeros(4, 6)
                sd[f"l{{l}}r{{r}}"] = ShardedTensor.from_rank_offsets("decoder.layers.self_attention.linear_qkv.weight", t,
                                                                    (0, l, 2), (1, r, 2), prepend_axis_num=1)
        sd["emb"] = ShardedTensor.from_rank_offsets("embedding.word_embeddings.weight", emb.clone() if fill else torch.zeros(10, 4))
        sd["obj"] = ShardedObject("decoder.final._extra_state", {{"v": 7}} if fill else None, (1,), (0,))
        return sd
    if mode == "save":
        dc.save(build(True), path)
    else:
        out = dc.load(build(False), path)
        ok = all(torch.equal(out[f"l{{l}}r{{r}}"], w[l, 4 * r:4 * r + 4]) for l in range(2) for r in range(2))
        print("REF_LOAD_OK" if ok and torch.equal(out["emb"], emb) and out["obj"] == {{"v": 7}} else "REF_LOAD_BAD", out["obj"])
''')


def _run_ref(mode, path, port):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    return subprocess.run([sys.executable, "-c", _REF_SNIPPET.format(ref=REF, port=port), mode, path], capture_output=True,
                          text=True, timeout=600, env=env, cwd=ROOT)


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "megatron", "core", "dist_checkpointing")),
                    reason="baseline/_ref (the unmodified reference) is not installed")
def test_interchange_with_the_unmodified_reference(tmp_path):
    """A checkpoint written by the reference's own dist_checkpointing loads here under a different sharding, and a
    checkpoint written here loads through the reference's ``dist_checkpointing.load``."""
    w = torch.arange(96, dtype=torch.float32).view(2, 8, 6)
    emb = torch.arange(40, dtype=torch.float32).view(10, 4)
    # (a) files produced by PyTorch's own DCP writer -- the storage layer the reference's saver sits on (its own
    #     FileSystemWriterAsync does not run against torch 2.11: ``_write_item() missing 'serialization_format'``) --
    #     read here under a different sharding: one shard per layer, embedding split in two
    import torch.distributed.checkpoint as dcp

    ref_dir = str(tmp_path / "by_dcp")
    dcp.save({"decoder.layers.self_attention.linear_qkv.weight": w.clone(), "embedding.word_embeddings.weight": emb.clone()},
             checkpoint_id=ref_dir, no_dist=True)
    mine = {f"l{l}": _st("decoder.layers.self_attention.linear_qkv.weight", torch.zeros(8, 6), pre=[(0, l, 2)]) for l in range(2)}
    mine.update({f"e{r}": _st("embedding.word_embeddings.weight", torch.zeros(5, 4), offsets=[(0, r, 2)]) for r in range(2)})
    td.load(mine, ref_dir)
    assert all(torch.equal(mine[f"l{l}"].data, w[l]) for l in range(2))
    assert torch.equal(torch.cat([mine["e0"].data, mine["e1"].data]), emb)
    # (b) written here, read by the unmodified reference's dist_checkpointing.load (4 TP x layer shards + object)
    own_dir = str(tmp_path / "by_us")
    ours = {f"l{l}": _st("decoder.layers.self_attention.linear_qkv.weight", w[l].clone(), pre=[(0, l, 2)]) for l in range(2)}
    ours["emb"] = _st("embedding.word_embeddings.weight", emb.clone())
    ours["obj"] = ShardedObject("decoder.final._extra_state", {"v": 7})
    td.save(ours, own_dir, {"iteration": 1}).finalize()
    p = _run_ref("load", own_dir, 29633)
    assert p.returncode == 0 and "REF_LOAD_OK" in p.stdout, (p.stdout[-1500:], p.stderr[-3000:])


def test_async_writer_process_defers_visibility(tmp_path):
    """--async-save: the files are written by a persistent worker process; .metadata appears only at finalize()."""
    ssd = {"w": _st("a.weight", torch.randn(64, 64))}
    h = td.save(ssd, str(tmp_path), {"iteration": 9}, async_save=True)
    assert not os.path.exists(tmp_path / ".metadata")
    h.finalize()
    assert os.path.exists(tmp_path / ".metadata") and os.path.exists(tmp_path / "__0_0.distcp")
    back = _st("a.weight", torch.zeros(64, 64))
    td.load({"w": back}, str(tmp_path))
    assert torch.equal(back.data, ssd["w"].data)
    td.shutdown_async_writer()


def test_dist_checkpointing_entry_points_strictness_metadata_removal_and_validation(tmp_path):
    """The reference's ``dist_checkpointing`` surface: strict modes of ``load``, metadata-only and whole-tensor loads, dropping a
    key family (optimizer state) from an existing checkpoint, non-persistent local objects, sharding validation."""
    import megatronapp_b200.ckpt as dc
    from megatronapp_b200.ckpt import CheckpointingException, LocalNonpersistentObject, StrictHandling

    w = torch.arange(48, dtype=torch.float32).view(8, 6)
    ssd = {"w#0": _st("blk.weight", w[:4], offsets=[(0, 0, 2)]), "w#1": _st("blk.weight", w[4:], offsets=[(0, 1, 2)]),
           "m": _st("optimizer.state.exp_avg.blk.weight", torch.ones(8, 6)), "obj": ShardedObject("blk._extra_state", {"a": 1}),
           "rng": LocalNonpersistentObject({"seed": 7})}
    dc.save(ssd, str(tmp_path), {"iteration": 3}).finalize()
    assert dc.check_is_distributed_checkpoint(str(tmp_path)) and dc.load_common_state_dict(str(tmp_path))["iteration"] == 3
    meta = dc.load_tensors_metadata(str(tmp_path))
    assert set(meta) == {"blk.weight", "optimizer.state.exp_avg.blk.weight"} and meta["blk.weight"].global_shape == (8, 6) and meta["blk.weight"].data is None
    assert "blk._extra_state/shard_0_1" in dc.load_sharded_metadata(str(tmp_path)) or any("_extra_state" in k for k in dc.load_sharded_metadata(str(tmp_path)))
    plain = dc.load_plain_tensors(str(tmp_path))
    assert torch.equal(plain["blk.weight"], w)
    # strictness
    want = {"w": _st("blk.weight", torch.zeros(8, 6)), "gone": _st("blk.bias", torch.zeros(6)), "rng": LocalNonpersistentObject({"seed": 9})}
    out, missing, unexpected = dc.load(want, str(tmp_path), StrictHandling.RETURN_ALL)
    assert torch.equal(out["w"], w) and out["rng"] == {"seed": 9} and missing == ["blk.bias"] and "optimizer.state.exp_avg.blk.weight" in unexpected
    with pytest.raises(CheckpointingException, match="missing"):
        dc.load(want, str(tmp_path), StrictHandling.RAISE_ALL)
    with pytest.raises(CheckpointingException, match="unexpected"):
        dc.load({"w": _st("blk.weight", torch.zeros(8, 6))}, str(tmp_path), StrictHandling.RAISE_UNEXPECTED)
    with pytest.raises(KeyError):
        dc.load(want, str(tmp_path), StrictHandling.ASSUME_OK_UNEXPECTED)
    # release checkpoint: optimizer state dropped from the metadata, the model still loads
    gone = dc.remove_sharded_tensors(str(tmp_path), "optimizer.")
    assert gone == ["optimizer.state.exp_avg.blk.weight"] and set(dc.load_tensors_metadata(str(tmp_path))) == {"blk.weight"}
    full = dc.load({"w": _st("blk.weight", torch.zeros(8, 6)), "obj": ShardedObject("blk._extra_state", None)}, str(tmp_path),
                   StrictHandling.RAISE_ALL)
    assert torch.equal(full["w"], w) and full["obj"] == {"a": 1}
    # validation: gap, overlap, out-of-bounds, flattened ranges
    dc.validate_sharding_integrity(ssd)
    with pytest.raises(CheckpointingException, match="cover 24 of 48"):
        dc.validate_sharding_integrity({"w#0": ssd["w#0"]})
    with pytest.raises(CheckpointingException, match="overlap"):
        dc.validate_sharding_integrity({"a": ssd["w#0"], "b": _st("blk.weight", w[:4], offsets=[(0, 0, 2)])})
    flat = {"a": _st("p", torch.zeros(10), flat=(0, 6)), "b": _st("p", torch.zeros(10), flat=(6, 10))}
    dc.validate_sharding_integrity(flat)
    with pytest.raises(CheckpointingException, match="cover 9 of 10"):
        dc.validate_sharding_integrity({"a": flat["a"], "b": _st("p", torch.zeros(10), flat=(7, 10))})
    with pytest.raises(CheckpointingException):
        dc.save({"w#0": ssd["w#0"]}, str(tmp_path / "bad"))
