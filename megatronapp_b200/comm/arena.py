"""Symmetric-memory arena over CUDA IPC.

Every rank of a group ``cudaMalloc``s one arena (native ``_C.arena_alloc``), exports it with
``cudaIpcGetMemHandle`` and maps every peer's arena, so the fused kernels
(``gemm_ag`` / ``gemm_rs`` / ``bucket_reduce_scatter`` / ``a2a_scatter_rows``) can issue plain
``ld/st`` to peer HBM over NVLink 5 / NVSwitch.  Regions are carved by a deterministic bump
allocator, so the same name maps to the same offset on every rank ("symmetric").

Synchronisation contract used by all kernels: per-region *epoch* counters (monotonic, one per
double-buffer slot); writers publish the epoch with ``st.release.sys``, readers poll with
``ld.acquire.sys``.  Alternating two slots per region makes reuse safe without a barrier:
a rank can be at most one collective ahead of any peer because every fused kernel needs data
from all peers before it completes.

This replaces the reference's userbuffer / shm plumbing (``initialize.py:236-292``,
``shm_tensor_new_rdma*.cpp``) for the single-node NVSwitch case.
"""
from __future__ import annotations

import os

from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

from .. import ops

_ALIGN = 1024


@dataclass
class Region:
    name: str
    offset: int
    nbytes: int
    epoch: int = 0


class SymmetricArena:
    def __init__(self, group, nbytes: int, device: Optional[torch.device] = None):
        if not torch.cuda.is_available():
            raise RuntimeError("SymmetricArena needs CUDA")
        self.group = group
        self.size = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.nbytes = (nbytes + _ALIGN - 1) // _ALIGN * _ALIGN
        ext = ops.ext()
        self.local_ptr, handle = ext.arena_alloc(self.nbytes)
        handles: List[Optional[bytes]] = [None] * self.size
        dist.all_gather_object(handles, handle, group=group)
        self.peer_ptrs: List[int] = []
        for r, h in enumerate(handles):
            self.peer_ptrs.append(self.local_ptr if r == self.rank else ext.arena_open(h))
        self._cursor = 0
        self.regions: Dict[str, Region] = {}
        dist.barrier(group=group)

    # -- allocation -----------------------------------------------------------
    def region(self, name: str, nbytes: int) -> Region:
        reg = self.regions.get(name)
        if reg is not None:
            if reg.nbytes < nbytes:
                raise RuntimeError(f"arena region {name!r} was sized {reg.nbytes} < requested {nbytes}")
            return reg
        nbytes = (nbytes + _ALIGN - 1) // _ALIGN * _ALIGN
        if self._cursor + nbytes > self.nbytes:
            raise RuntimeError(
                f"symmetric arena exhausted: {self._cursor}+{nbytes} > {self.nbytes} (region {name!r})")
        reg = Region(name, self._cursor, nbytes)
        self._cursor += nbytes
        self.regions[name] = reg
        return reg

    def ptrs(self, reg: Region, extra_offset: int = 0) -> List[int]:
        return [p + reg.offset + extra_offset for p in self.peer_ptrs]

    def local(self, reg: Region, shape, dtype) -> torch.Tensor:
        return ops.ext().tensor_from_ptr(self.local_ptr + reg.offset, list(shape), dtype, self.device.index)

    def next_epoch(self, reg: Region) -> int:
        reg.epoch += 1
        return reg.epoch

    def barrier(self, slot: int = 0) -> None:
        """Device-side barrier over the group (one tiny kernel, no host sync)."""
        reg = self.region("__barrier__", 4 * 8 * 16)
        ep = self.next_epoch(reg)
        ops.ext().flags_barrier(self.ptrs(reg), self.rank, self.size, slot, ep)

    def close(self) -> None:
        ext = ops.ext()
        torch.cuda.synchronize()
        for r, p in enumerate(self.peer_ptrs):
            if r != self.rank:
                ext.arena_close(p)
        ext.arena_free(self.local_ptr)
        self.peer_ptrs = []


_ARENAS: Dict[str, SymmetricArena] = {}


def get_arena(kind: str, group=None, nbytes: int = 0) -> Optional[SymmetricArena]:
    """Per-group arena registry (``kind`` in tp / dp / ep)."""
    a = _ARENAS.get(kind)
    if a is None and group is not None and nbytes > 0:
        a = SymmetricArena(group, nbytes)
        _ARENAS[kind] = a
    return a


def destroy_arenas() -> None:
    for a in _ARENAS.values():
        try:
            a.close()
        except Exception:
            pass
    _ARENAS.clear()


class TPFusedComm:
    """Book-keeping for the fused AG->GEMM / GEMM->RS kernels of one TP group.

    Two slots per direction (alternating per call); gather buffers sized for the largest
    ``[tokens, hidden]`` activation, RS staging for the largest ``[tokens, out]`` partial.
    """

    def __init__(self, arena: SymmetricArena, max_tokens: int, max_in: int, max_out: int,
                 n_comm_ctas: int = 8):
        self.arena = arena
        # comm CTAs push the local shard to tp-1 peers with TMA bulk copies; the bytes in flight per CTA are
        # fixed (6 x 32 KB), so the count scales with the number of destinations (override: B200_AG_COMM_CTAS)
        env = os.environ.get("B200_AG_COMM_CTAS")
        self.n_comm_ctas = int(env) if env else (n_comm_ctas if arena.size <= 2 else 16)
        # GEMM->RS: pushed partial tiles + reduction folded into the epilogue of the owner's own tiles (cta_group::2
        # kernel, no serialised reduce tail).  On by default; B200_FUSED_RS=0 / bench.py --no-fused-rs falls back to
        # tcgen05 GEMM + NCCL reduce-scatter (the in-tree baseline).
        self.use_fused_rs = os.environ.get("B200_FUSED_RS", "1") == "1"
        # column-parallel wgrad with the activation re-gather fused along its reduction axis.  Opt-in (B200_FUSED_AG_WGRAD=1):
        # measured at TP4 (profiles/r2_fused_comm_4gpu_agwgrad.json) it is exact but 0.42 ms vs 0.31 ms for the NCCL
        # all-gather overlapped with the dgrad kernel + plain wgrad GEMM, and the 6.7B step is 4 % slower with it.
        self.use_fused_ag_wgrad = os.environ.get("B200_FUSED_AG_WGRAD", "0") == "1"
        # 0 = auto (cta_group::2 kernel when the shape allows), 128 / 256 = force the 1-CTA kernel (A/B measurements)
        self.block_n = int(os.environ.get("B200_FUSED_BLOCK_N", "0"))
        self.calls = {"ag": 0, "rs": 0, "embed": 0}
        # K14: vocab-parallel embedding lookup fused with the sequence-parallel scatter.  Kernel-level it is bit-exact and
        # 2.6x faster than F.embedding + mask + NCCL reduce-scatter (profiles/r2_bucket_rs_embed_2gpu.json), but the in-model
        # path (autograd backward through the tied embedding) has not run a training step on hardware yet, and it saves
        # ~0.1 ms per step: opt-in (B200_FUSED_EMBED=1) until it has.
        self.use_fused_embed = os.environ.get("B200_FUSED_EMBED", "0") == "1"
        self.embed_ctas = 32
        self._embed = None
        ext = ops.ext()
        nflags = max(ext.gemm_flags_needed(max_tokens, max(max_in, max_out), arena.size), 4096)
        self.ag = [arena.region(f"tp_ag{i}", max_tokens * max_in * 2) for i in range(2)]
        self.ag_flags = [arena.region(f"tp_ag_flags{i}", nflags * 4) for i in range(2)]
        self.rs = [arena.region(f"tp_rs{i}", max_tokens * max_out * 2) for i in range(2)]
        self.rs_flags = [arena.region(f"tp_rs_flags{i}", nflags * 4) for i in range(2)]

    def embed_fits(self, n_tokens: int, hidden: int) -> bool:
        need = n_tokens * hidden * 2 // self.arena.size
        if self._embed is None:
            try:
                self._embed = ([self.arena.region(f"tp_embed{i}", need) for i in range(2)],
                               [self.arena.region(f"tp_embed_flags{i}", self.arena.size * self.embed_ctas * 4) for i in range(2)])
            except RuntimeError:
                self.use_fused_embed = False
                return False
        return self._embed[0][0].nbytes >= need

    def embed(self, tokens: torch.Tensor, weight: torch.Tensor, vocab_start: int) -> torch.Tensor:
        """tokens [b, s] int64, weight [V / tp, h] -> [s / tp, b, h]; rows arrive from their owners over NVLink."""
        a = self.arena
        slot = self.calls["embed"] % 2
        self.calls["embed"] += 1
        out_reg, flag_reg = self._embed[0][slot], self._embed[1][slot]
        ep = a.next_epoch(flag_reg)
        ops.ext().embed_scatter(tokens, weight, a.ptrs(out_reg), a.ptrs(flag_reg), int(vocab_start), a.rank, a.size, ep,
                                self.embed_ctas)
        b, s = tokens.shape
        # the slot is rewritten two calls later: hand autograd its own copy
        return a.local(out_reg, (s // a.size, b, weight.shape[1]), weight.dtype).clone()

    def ag_gemm(self, x_local: torch.Tensor, weight: torch.Tensor, bias=None, b_mn: bool = False,
                act: int = 0, out: Optional[torch.Tensor] = None,
                out_pre: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """``AG(x_local) @ W^T``; returns (out, gathered_x_view).  The gathered view aliases the
        arena slot and is only valid until the slot is reused two calls later."""
        a = self.arena
        slot = self.calls["ag"] & 1
        self.calls["ag"] += 1
        reg, freg = self.ag[slot], self.ag_flags[slot]
        epoch = a.next_epoch(freg)
        M = x_local.shape[0] * a.size
        K = x_local.shape[1]
        N = weight.shape[1] if b_mn else weight.shape[0]
        if out is None:
            out = torch.empty(M, N, dtype=x_local.dtype, device=x_local.device)
        ops.ext().gemm_ag(x_local, a.local_ptr + reg.offset, weight, out, b_mn, bias, act,
                          a.ptrs(reg), a.ptrs(freg), a.rank, a.size, epoch, self.n_comm_ctas, self.block_n, out_pre)
        gathered = a.local(reg, (M, K), x_local.dtype)
        return out, gathered

    def ag_wgrad(self, dy: torch.Tensor, x_local: torch.Tensor, out32: torch.Tensor, accumulate: bool) -> None:
        """``out32 (+)= dy^T @ all_gather(x_local)``: the column-parallel weight gradient with the re-gather of the
        sequence-parallel activation fused into the GEMM's reduction loop (no NCCL all-gather next to the dgrad kernel)."""
        a = self.arena
        slot = self.calls["ag"] & 1
        self.calls["ag"] += 1
        reg, freg = self.ag[slot], self.ag_flags[slot]
        epoch = a.next_epoch(freg)
        ops.ext().gemm_ag_wgrad(dy, x_local, a.local_ptr + reg.offset, out32, accumulate, a.ptrs(reg), a.ptrs(freg), a.rank,
                                a.size, epoch, self.n_comm_ctas)

    def gemm_rs(self, x: torch.Tensor, weight: torch.Tensor, bias=None, b_mn: bool = False) -> torch.Tensor:
        """``reduce_scatter(x @ W^T)`` along rows."""
        a = self.arena
        slot = self.calls["rs"] & 1
        self.calls["rs"] += 1
        reg, freg = self.rs[slot], self.rs_flags[slot]
        epoch = a.next_epoch(freg)
        M = x.shape[0]
        N = weight.shape[1] if b_mn else weight.shape[0]
        out = torch.empty(M // a.size, N, dtype=x.dtype, device=x.device)
        ops.ext().gemm_rs(x, weight, out, b_mn, bias, a.ptrs(reg), a.ptrs(freg), a.rank, a.size, epoch, self.block_n)
        return out
