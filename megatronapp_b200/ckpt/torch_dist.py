"""``torch_dist`` checkpoints in the reference's on-disk layout (PyTorch Distributed Checkpoint).

A checkpoint directory written by the reference (``dist_checkpointing/strategies/torch.py``: ``TorchDistSaveShardedStrategy``
over ``torch.distributed.checkpoint``) contains::

    .metadata                 pickled ``torch.distributed.checkpoint.metadata.Metadata``: per key the global size and the
                              list of stored chunks (offsets / sizes), and ``storage_data``: chunk -> (file, byte offset, length)
    __{rank}_{n}.distcp       concatenated ``torch.save`` records, one per stored chunk / object
    common.pt                 non-sharded state (args, iteration, rng, scheduler ...)       (strategies/common.py)
    metadata.json             {"sharded_backend": "torch_dist", "sharded_backend_version": 1, "common_backend": "torch", ...}

and its keys follow ``ShardedTensor.key`` (mapping.py:79-89): homogeneous transformer layers are STACKED under one key
(``decoder.layers.self_attention.linear_qkv.weight`` with a leading ``num_layers`` axis, transformer_block.py:600-620),
gated-MLP ``linear_fc1`` is stored gate-block-then-up-block (mlp.py ``apply_swiglu_sharded_factory``), module
``_extra_state`` entries are byte objects ``{key}/shard_{offset}_{shape}``, distributed-optimizer state is stored in model
space as N-D "flattened range" tensors (reformulated global shape ``[fragments..., local_numel]``, torch.py:131-250).

This module writes and reads exactly that layout with its own (communication-free) planner:

* save: every rank appends the shards it is the designated writer of to ``__{rank}_0.distcp`` -- the writer of a
  replicated shard is picked by a stable hash over the replicas ("fully parallel" save, fully_parallel.py:50-120, without the
  exchange round) -- the per-rank chunk tables are gathered and rank 0 pickles the ``Metadata``.  The file I/O can run in a
  persistent writer *process* (``AsyncWriter``); the checkpoint only becomes visible (``.metadata`` + tracker) in the
  finalize step after every rank's writer finished (async_utils.py:90-230).
* load: every requested local shard is assembled from the intersecting stored chunks, whatever parallel layout wrote
  them (TP / PP / EP / DP resharding is a property of the format); flattened-range tensors are re-assembled through their
  N-D boxes so optimizer state reshards too.
"""
from __future__ import annotations

import io
import json
import os
import pickle
import re
import zlib
from dataclasses import replace
from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist
from torch.distributed.checkpoint.filesystem import _StorageInfo
from torch.distributed.checkpoint.metadata import (BytesStorageMetadata, ChunkStorageMetadata, Metadata, MetadataIndex,
                                                   TensorProperties, TensorStorageMetadata)

from .sharded import ShardedObject, ShardedTensor

METADATA_JSON = {"sharded_backend": "torch_dist", "sharded_backend_version": 1, "common_backend": "torch",
                 "common_backend_version": 1}
_LAYER_RE = re.compile(r"^(.*\blayers\.)(\d+)\.(.*)$")


def _rank() -> int:
    return dist.get_rank() if dist.is_initialized() else 0


def _world() -> int:
    return dist.get_world_size() if dist.is_initialized() else 1


# --------------------------------------------------------------------------------------------------- key layout
def stack_homogeneous_layers(ssd: Dict[str, object], num_layers: int) -> Dict[str, object]:
    """``...layers.<g>.<rest>`` -> ``...layers.<rest>`` with a prepended layer axis (offset g of num_layers), the reference's
    default layout for models whose layers are all alike (transformer_block.py:600-620).  Dict keys stay unique."""
    out = {}
    for k, v in ssd.items():
        m = _LAYER_RE.match(v.key) if isinstance(v, (ShardedTensor, ShardedObject)) else None
        if m is None:
            out[k] = v
            continue
        g = int(m.group(2))
        key = m.group(1) + m.group(3)
        if isinstance(v, ShardedTensor):
            out[k] = replace(v, key=key, global_shape=(num_layers,) + tuple(v.global_shape),
                             global_offset=(g,) + tuple(v.global_offset),
                             axis_fragmentations=(num_layers,) + tuple(v.axis_fragmentations or (1,) * len(v.global_shape)),
                             prepend_axis_num=v.prepend_axis_num + 1)
        else:
            out[k] = ShardedObject(key, v.data, v.replica_id, global_shape=(num_layers,) + tuple(getattr(v, "global_shape", ()) or ()),
                                   global_offset=(g,) + tuple(getattr(v, "global_offset", ()) or ()))
    return out


def sharded_object_key(o: ShardedObject) -> str:
    gs = tuple(getattr(o, "global_shape", None) or (1,))
    go = tuple(getattr(o, "global_offset", None) or (0,))
    return f"{o.key}/shard_{'.'.join(map(str, go))}_{'.'.join(map(str, gs))}"


# --------------------------------------------------------------------------------------------------- chunk geometry
def _is_writer(replica_id, fqn: str, offsets: Tuple[int, ...], replica_count: Optional[int] = None) -> bool:
    """One writer per distinct shard.  ``replica_id`` is ``(pp-like, tp-like, dp)``: a shard is replicated over the last
    component (data-parallel copies) when the leading ones are 0; the copy whose dp index equals a stable hash of the shard
    writes it, so the write load spreads over the data-parallel group instead of landing on dp rank 0."""
    rid = replica_id if isinstance(replica_id, (tuple, list)) else (replica_id,)
    if any(int(x) != 0 for x in rid[:-1]):
        return False
    dp = int(rid[-1])
    n_dp = int(replica_count) if replica_count else 1     # unknown replica-group size: the main replica writes
    if n_dp <= 1:
        return dp == 0
    h = zlib.crc32((fqn + "@" + ",".join(map(str, offsets))).encode())
    return dp == h % n_dp


def _reformulated(st: ShardedTensor):
    """N-D flattened tensors are stored as ``[fragments..., local_numel]`` with the last axis cut by ``flattened_range``
    (torch.py:150-160 ``nd_flattened_tensor_reformulated_global_shape``)."""
    frags = tuple(st.axis_fragmentations or (1,) * len(st.global_shape))
    local_numel = 1
    for n in st.local_shape:
        local_numel *= int(n)
    gshape = frags + (local_numel,)
    chunk = tuple(int(o) // max(1, int(s)) for o, s in zip(st.global_offset[st.prepend_axis_num:], st.local_shape))
    chunk = tuple(int(o) for o in st.global_offset[:st.prepend_axis_num]) + chunk
    fr = st.flattened_range
    offsets = chunk + (int(fr.start),)
    sizes = (1,) * len(chunk) + (int(fr.stop - fr.start),)
    return gshape, offsets, sizes


def _chunk_view(st: ShardedTensor):
    """(fqn, global size, chunk offsets, chunk sizes, tensor to store) of one local shard."""
    data = st.data
    if hasattr(data, "full_tensor"):
        data = data.full_tensor()
    if st.flattened_range is not None:
        gshape, offsets, sizes = _reformulated(st)
        return st.key, gshape, offsets, sizes, data.detach().reshape(sizes)
    sizes = (1,) * st.prepend_axis_num + tuple(int(x) for x in st.local_shape)
    return st.key, tuple(int(x) for x in st.global_shape), tuple(int(x) for x in st.global_offset), sizes, data.detach().reshape(sizes)


# --------------------------------------------------------------------------------------------------- async writer process
def _writer_main(q_in, q_out):
    """Persistent writer process: jobs are (job id, path, [(tag, tensor-or-bytes)]); replies (job id, [(tag, offset, length)])."""
    torch.set_num_threads(1)
    while True:
        job = q_in.get()
        if job is None:
            return
        jid, path, items = job
        try:
            q_out.put((jid, _write_records(path, items), None))
        except BaseException as e:  # noqa: BLE001
            q_out.put((jid, None, repr(e)))


def _write_records(path: str, items) -> List[Tuple[object, int, int]]:
    out = []
    tmp = path + ".tmp"
    with open(tmp, "wb") as f:
        for tag, payload in items:
            pos = f.tell()
            if isinstance(payload, (bytes, bytearray)):
                f.write(payload)
            else:
                torch.save(payload, f)
            out.append((tag, pos, f.tell() - pos))
        f.flush()
        os.fsync(f.fileno())
    os.replace(tmp, path)
    return out


class AsyncWriter:
    """A long-lived writer process (reference ``PersistentAsyncCaller``, async_utils.py:300-480): the training process stages
    tensors in shared CPU memory and returns immediately; ``poll`` / ``wait`` collect the byte offsets of finished jobs."""

    def __init__(self):
        import torch.multiprocessing as mp

        ctx = mp.get_context("spawn")
        self.q_in, self.q_out = ctx.Queue(), ctx.Queue()
        self.proc = ctx.Process(target=_writer_main, args=(self.q_in, self.q_out), daemon=True)
        self.proc.start()
        self._next, self._done = 0, {}

    def submit(self, path: str, items) -> int:
        jid = self._next
        self._next += 1
        shared = [(tag, p if isinstance(p, (bytes, bytearray)) else p.share_memory_()) for tag, p in items]
        self.q_in.put((jid, path, shared))
        return jid

    def poll(self):
        import queue

        while True:
            try:
                jid, res, err = self.q_out.get_nowait()
            except queue.Empty:
                return
            self._done[jid] = (res, err)

    def ready(self, jid: int) -> bool:
        self.poll()
        return jid in self._done

    def wait(self, jid: int):
        while jid not in self._done:
            j, res, err = self.q_out.get()
            self._done[j] = (res, err)
        res, err = self._done.pop(jid)
        if err is not None:
            raise RuntimeError(f"async checkpoint writer failed: {err}")
        return res

    def close(self):
        if self.proc.is_alive():
            self.q_in.put(None)
            self.proc.join(timeout=30)


_WRITER: List[Optional[AsyncWriter]] = [None]


def get_async_writer() -> AsyncWriter:
    if _WRITER[0] is None or not _WRITER[0].proc.is_alive():
        _WRITER[0] = AsyncWriter()
    return _WRITER[0]


def shutdown_async_writer() -> None:
    if _WRITER[0] is not None:
        _WRITER[0].close()
        _WRITER[0] = None


# --------------------------------------------------------------------------------------------------- save
class PendingSave:
    """Handle of a (possibly asynchronous) save: ``finalize()`` gathers the chunk tables, writes ``.metadata`` /
    ``common.pt`` / ``metadata.json`` on rank 0 and synchronises the ranks -- only then is the checkpoint complete."""

    def __init__(self, checkpoint_dir, fname, tensor_meta, bytes_keys, records, job, common_state, mcore_data):
        self.dir, self.fname = checkpoint_dir, fname
        self.tensor_meta, self.bytes_keys, self.records, self.job = tensor_meta, bytes_keys, records, job
        self.common_state, self.mcore_data = common_state, mcore_data
        self.finalized = False

    def ready(self) -> bool:
        return self.job is None or get_async_writer().ready(self.job)

    def finalize(self) -> None:
        if self.finalized:
            return
        if self.job is not None:
            self.records = get_async_writer().wait(self.job)
        storage = {}
        for (fqn, offsets), pos, length in self.records:
            idx = MetadataIndex(fqn, torch.Size(offsets)) if offsets is not None else MetadataIndex(fqn)
            storage[idx] = _StorageInfo(self.fname, pos, length)
        part = (self.tensor_meta, self.bytes_keys, storage, self.mcore_data)
        if dist.is_initialized() and _world() > 1:
            parts = [None] * _world()
            dist.all_gather_object(parts, part)
        else:
            parts = [part]
        if _rank() == 0:
            sdm: Dict[str, object] = {}
            storage_all = {}
            mcore = {}
            for tmeta, bkeys, stor, mc in parts:
                for fqn, (props, size, chunks) in tmeta.items():
                    cur = sdm.get(fqn)
                    if cur is None:
                        cur = sdm[fqn] = TensorStorageMetadata(props, torch.Size(size), [])
                    elif tuple(cur.size) != tuple(size):
                        raise ValueError(f"inconsistent global shape for {fqn}: {tuple(cur.size)} vs {tuple(size)}")
                    have = {tuple(c.offsets) for c in cur.chunks}
                    for off, sz in chunks:
                        if tuple(off) not in have:
                            cur.chunks.append(ChunkStorageMetadata(torch.Size(off), torch.Size(sz)))
                for k in bkeys:
                    sdm[k] = BytesStorageMetadata()
                storage_all.update(stor)
                mcore.update(mc)
            md = Metadata(state_dict_metadata=sdm, planner_data=None, storage_data=storage_all)
            md.mcore_data = mcore            # reference torch.py:495-500 attaches the reformulation table the same way
            tmp = os.path.join(self.dir, ".metadata.tmp")
            with open(tmp, "wb") as f:
                pickle.dump(md, f)
                f.flush()
                os.fsync(f.fileno())
            os.replace(tmp, os.path.join(self.dir, ".metadata"))
            if self.common_state is not None:
                torch.save(self.common_state, os.path.join(self.dir, "common.pt"))
            with open(os.path.join(self.dir, "metadata.json"), "w") as f:
                json.dump(METADATA_JSON, f)
        if dist.is_initialized() and _world() > 1:
            dist.barrier()
        self.finalized = True


def save(sharded_state_dict: Dict[str, object], checkpoint_dir: str, common_state: Optional[dict] = None,
         async_save: bool = False) -> PendingSave:
    """Write this rank's shards; returns a handle whose ``finalize()`` completes the checkpoint (call it right away for a
    synchronous save, or from the training loop's ``maybe_finalize_async_save`` for ``--async-save``)."""
    os.makedirs(checkpoint_dir, exist_ok=True)
    rank = _rank()
    fname = f"__{rank}_0.distcp"
    tensor_meta: Dict[str, tuple] = {}
    bytes_keys: List[str] = []
    items = []
    mcore_data = {}
    for v in sharded_state_dict.values():
        if isinstance(v, ShardedTensor):
            fqn, gsize, offsets, sizes, t = _chunk_view(v)
            props = TensorProperties(dtype=v.dtype)
            ent = tensor_meta.setdefault(fqn, (props, gsize, []))
            if tuple(ent[1]) != tuple(gsize):
                raise ValueError(f"inconsistent global shape for {fqn}")
            ent[2].append((offsets, sizes))
            if v.flattened_range is not None:
                mcore_data[fqn] = {"nd_reformulated_orig_global_shape": tuple(int(x) for x in v.global_shape)}
            if _is_writer(v.replica_id, fqn, offsets, v.replica_count):
                items.append(((fqn, offsets), t.to("cpu", copy=True).contiguous()))
        elif isinstance(v, ShardedObject):
            k = sharded_object_key(v)
            bytes_keys.append(k)
            rid = v.replica_id if isinstance(v.replica_id, (tuple, list)) else (v.replica_id,)
            if all(int(x) == 0 for x in rid):
                buf = io.BytesIO()
                torch.save([v.data], buf)        # the reference packs the objects of a key in a list (torch.py:341-346)
                items.append(((k, None), buf.getvalue()))
    path = os.path.join(checkpoint_dir, fname)
    if async_save:
        job = get_async_writer().submit(path, items)
        return PendingSave(checkpoint_dir, fname, tensor_meta, bytes_keys, None, job, common_state, mcore_data)
    records = _write_records(path, items)
    return PendingSave(checkpoint_dir, fname, tensor_meta, bytes_keys, records, None, common_state, mcore_data)


# --------------------------------------------------------------------------------------------------- load
def is_torch_dist_checkpoint(checkpoint_dir: str) -> bool:
    return os.path.exists(os.path.join(checkpoint_dir, ".metadata"))


def read_metadata(checkpoint_dir: str) -> Metadata:
    with open(os.path.join(checkpoint_dir, ".metadata"), "rb") as f:
        return pickle.load(f)


def load_common_state_dict(checkpoint_dir: str) -> dict:
    p = os.path.join(checkpoint_dir, "common.pt")
    return torch.load(p, map_location="cpu", weights_only=False) if os.path.exists(p) else {}


class _Reader:
    def __init__(self, checkpoint_dir: str, md: Metadata):
        self.dir, self.md = checkpoint_dir, md
        self.files: Dict[str, object] = {}
        self.by_fqn: Dict[str, List[Tuple[tuple, _StorageInfo]]] = {}
        for idx, info in (md.storage_data or {}).items():
            self.by_fqn.setdefault(idx.fqn, []).append((tuple(idx.offset) if idx.offset is not None else None, info))

    def _bytes(self, info: _StorageInfo) -> bytes:
        f = self.files.get(info.relative_path)
        if f is None:
            f = self.files[info.relative_path] = open(os.path.join(self.dir, info.relative_path), "rb")
        f.seek(info.offset)
        return f.read(info.length)

    def tensor(self, fqn: str, offsets: tuple) -> torch.Tensor:
        for off, info in self.by_fqn.get(fqn, ()):
            if off == tuple(offsets):
                return torch.load(io.BytesIO(self._bytes(info)), map_location="cpu", weights_only=False)
        raise KeyError(f"chunk {offsets} of {fqn} has no storage record")

    def obj(self, fqn: str):
        for off, info in self.by_fqn.get(fqn, ()):
            return torch.load(io.BytesIO(self._bytes(info)), map_location="cpu", weights_only=False)
        raise KeyError(fqn)

    def close(self):
        for f in self.files.values():
            f.close()


def _fill_box(dst: torch.Tensor, dst_off: Tuple[int, ...], rd: _Reader, fqn: str, tmeta: TensorStorageMetadata) -> int:
    """Copy every stored chunk's intersection with the box (dst_off, dst.shape) into dst; returns covered element count."""
    lo = tuple(dst_off)
    hi = tuple(o + s for o, s in zip(lo, dst.shape))
    covered = 0
    for ch in tmeta.chunks:
        so, ss = tuple(ch.offsets), tuple(ch.sizes)
        se = tuple(o + s for o, s in zip(so, ss))
        ilo = tuple(max(a, b) for a, b in zip(lo, so))
        ihi = tuple(min(a, b) for a, b in zip(hi, se))
        if any(a >= b for a, b in zip(ilo, ihi)):
            continue
        src = rd.tensor(fqn, so)
        src = src.reshape(ss)
        src_sl = tuple(slice(a - o, b - o) for a, b, o in zip(ilo, ihi, so))
        dst_sl = tuple(slice(a - o, b - o) for a, b, o in zip(ilo, ihi, lo))
        dst[dst_sl] = src[src_sl].to(dst.dtype)
        n = 1
        for a, b in zip(ilo, ihi):
            n *= b - a
        covered += n
    return covered


def _load_flattened(v: ShardedTensor, rd: _Reader, md: Metadata, tmeta: TensorStorageMetadata) -> torch.Tensor:
    """A flattened-range shard under a possibly different layout: rebuild this rank's N-D box from the stored flat pieces
    (each is a run inside the row-major flattening of the box of the rank that wrote it), then take the requested run."""
    orig = (getattr(md, "mcore_data", None) or {}).get(v.key, {}).get("nd_reformulated_orig_global_shape")
    nd = len(v.global_shape)
    if orig is None or len(tmeta.size) != nd + 1:
        raise ValueError(f"{v.key}: checkpoint entry is not an N-D flattened tensor")
    frags = tuple(int(x) for x in tmeta.size[:-1])
    src_local = tuple(int(g) // f for g, f in zip(orig, frags))
    box = torch.zeros(tuple(int(x) for x in v.local_shape), dtype=v.dtype)
    have = torch.zeros(box.shape, dtype=torch.bool)
    pre = v.prepend_axis_num
    lo = tuple(int(x) for x in v.global_offset)
    lo_nd = lo[:pre] + lo[pre:]
    shape_nd = (1,) * pre + tuple(int(x) for x in v.local_shape)
    hi_nd = tuple(o + s for o, s in zip(lo_nd, shape_nd))
    for ch in tmeta.chunks:
        cidx, fstart = tuple(ch.offsets[:-1]), int(ch.offsets[-1])
        flen = int(ch.sizes[-1])
        s_lo = tuple(c * l for c, l in zip(cidx, src_local))
        s_hi = tuple(a + l for a, l in zip(s_lo, src_local))
        ilo = tuple(max(a, b) for a, b in zip(lo_nd, s_lo))
        ihi = tuple(min(a, b) for a, b in zip(hi_nd, s_hi))
        if any(a >= b for a, b in zip(ilo, ihi)):
            continue
        piece = rd.tensor(v.key, tuple(ch.offsets)).reshape(-1)
        full = torch.zeros(src_local, dtype=piece.dtype).reshape(-1)
        mask = torch.zeros(full.shape, dtype=torch.bool)
        full[fstart:fstart + flen] = piece
        mask[fstart:fstart + flen] = True
        full, mask = full.reshape(src_local), mask.reshape(src_local)
        s_sl = tuple(slice(a - o, b - o) for a, b, o in zip(ilo, ihi, s_lo))
        d_sl = tuple(slice(a - o, b - o) for a, b, o in zip(ilo, ihi, lo_nd))[pre:]
        sub, m = full[s_sl], mask[s_sl]
        if pre:
            sub, m = sub.reshape(sub.shape[pre:]), m.reshape(m.shape[pre:])
        cur = box[d_sl]
        box[d_sl] = torch.where(m, sub.to(box.dtype), cur)
        have[d_sl] |= m
    fr = v.flattened_range
    if not bool(have.reshape(-1)[fr.start:fr.stop].all()):
        raise ValueError(f"checkpoint does not cover the requested range of {v.key}")
    return box.reshape(-1)[fr.start:fr.stop].clone()


def load(sharded_state_dict: Dict[str, object], checkpoint_dir: str, strict: bool = True) -> Dict[str, object]:
    """Fill every ShardedTensor's ``data`` in place (ShardedObjects are returned) from a ``torch_dist`` checkpoint."""
    md = read_metadata(checkpoint_dir)
    rd = _Reader(checkpoint_dir, md)
    out, missing = {}, []
    try:
        for k, v in sharded_state_dict.items():
            if isinstance(v, ShardedObject):
                ok = sharded_object_key(v)
                if ok in md.state_dict_metadata:
                    val = rd.obj(ok)
                    out[k] = val[0] if isinstance(val, list) and len(val) == 1 else val
                elif strict:
                    missing.append(ok)
                continue
            if not isinstance(v, ShardedTensor):
                continue
            tmeta = md.state_dict_metadata.get(v.key)
            if tmeta is None or isinstance(tmeta, BytesStorageMetadata):
                missing.append(v.key)
                continue
            if v.flattened_range is not None:
                dst = _load_flattened(v, rd, md, tmeta)
            else:
                gshape = tuple(int(x) for x in tmeta.size)
                if gshape != tuple(int(x) for x in v.global_shape) and not v.allow_shape_mismatch:
                    raise ValueError(f"global shape mismatch for {v.key}: checkpoint {gshape} vs model {tuple(v.global_shape)}")
                shape = (1,) * v.prepend_axis_num + tuple(int(x) for x in v.local_shape)
                dst = torch.zeros(shape, dtype=v.dtype)
                covered = _fill_box(dst, tuple(int(x) for x in v.global_offset), rd, v.key, tmeta)
                if covered < dst.numel() and not v.allow_shape_mismatch:
                    raise ValueError(f"checkpoint chunks do not cover {v.key} (covered {covered} of {dst.numel()})")
                dst = dst.reshape(tuple(int(x) for x in v.local_shape))
            if v.data is not None and hasattr(v.data, "full_tensor"):
                from torch.distributed.tensor import distribute_tensor

                local = distribute_tensor(dst.to(v.data.device), v.data.device_mesh, v.data.placements)
                with torch.no_grad():
                    v.data.detach().to_local().copy_(local.to_local())
            elif v.data is not None:
                with torch.no_grad():
                    v.data.data.copy_(dst.reshape(v.data.shape).to(v.data.device))
            out[k] = dst
    finally:
        rd.close()
    if missing and strict:
        raise KeyError(f"keys missing from the checkpoint: {missing[:8]}{'...' if len(missing) > 8 else ''}")
    return out
