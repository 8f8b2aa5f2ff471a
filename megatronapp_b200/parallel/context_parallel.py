"""Context parallelism: the sequence is cut into ``2*cp`` chunks and rank ``r`` owns chunks
``(r, 2cp-1-r)`` so causal work is balanced.

Parity: reference ``core/utils.py get_batch_on_this_cp_rank :1704-1734``, rotary slicing
``rope_utils.get_pos_emb_on_this_cp_rank``, and the TE attention ``cp_comm_type`` modes
(``transformer_config.py:458-462``, ``extensions/transformer_engine.py:609-700``):

* ``a2a``        Ulysses: all-to-all heads <-> sequence, full-sequence attention on ``h/cp`` heads
                 with the native flash kernel, all-to-all back.
* ``all_gather`` K/V all-gathered along the sequence (reduce-scatter of dK/dV in backward),
                 local queries attend with a position mask.
* ``p2p``        K/V blocks travel around the CP ring (send to next / receive from previous,
                 ``cp - 1`` hops); attention runs on the assembled keys with the same mask.
* ``a2a+p2p``    hierarchical (``--hierarchical-context-parallel-sizes a p``, ``cp = a * p``): an all-to-all inside groups
                 of ``a`` consecutive CP ranks trades sequence for heads, then the K/V of the ``p`` outer groups travel around
                 a ring; without the sizes it degenerates to ``a2a`` over the whole CP group.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.distributed as dist

from . import state as ps


def _cp():
    g = ps.get_context_parallel_group()
    n = ps.get_context_parallel_world_size()
    return g, n, (ps.get_context_parallel_rank() if n > 1 else 0)


def chunk_ids(cp: int, rank: int):
    return rank, 2 * cp - 1 - rank


def local_positions(seq_len: int, cp: int, rank: int, device=None) -> torch.Tensor:
    """Global token positions held by ``rank`` (its two chunks, in local order)."""
    c = seq_len // (2 * cp)
    a, b = chunk_ids(cp, rank)
    return torch.cat([torch.arange(a * c, (a + 1) * c, device=device), torch.arange(b * c, (b + 1) * c, device=device)])


def gathered_order(seq_len: int, cp: int, device=None) -> torch.Tensor:
    """Global position of every row after concatenating the ranks' local sequences in rank order."""
    return torch.cat([local_positions(seq_len, cp, r, device) for r in range(cp)])


def slice_for_cp_rank(t: torch.Tensor, seq_dim: int, cp: Optional[int] = None, rank: Optional[int] = None) -> torch.Tensor:
    if cp is None:
        _, cp, rank = _cp()
    if cp == 1:
        return t
    s = t.shape[seq_dim]
    assert s % (2 * cp) == 0, f"sequence length {s} must divide by 2*cp={2 * cp}"
    idx = local_positions(s, cp, rank, t.device)
    return t.index_select(seq_dim, idx)


def get_batch_on_this_cp_rank(batch: Dict[str, Optional[torch.Tensor]]) -> Dict[str, Optional[torch.Tensor]]:
    _, cp, rank = _cp()
    if cp == 1:
        return batch
    out = {}
    for k, v in batch.items():
        if v is None:
            out[k] = None
        elif k == "attention_mask":
            out[k] = slice_for_cp_rank(v, 2, cp, rank)
        else:
            out[k] = slice_for_cp_rank(v, 1, cp, rank)
    return out


def get_pos_emb_on_this_cp_rank(pos_emb: torch.Tensor, seq_dim: int = 0) -> torch.Tensor:
    return slice_for_cp_rank(pos_emb, seq_dim)


# ------------------------------------------------------------------------------ collectives
class _GatherSeq(torch.autograd.Function):
    """All-gather along dim ``dim`` over the CP group; backward reduce-scatters the gradient."""

    @staticmethod
    def forward(ctx, x, dim, group):
        n = dist.get_world_size(group)
        ctx.dim, ctx.group, ctx.n = dim, group, n
        xs = [torch.empty_like(x) for _ in range(n)]
        dist.all_gather(xs, x.contiguous(), group=group)
        return torch.cat(xs, dim=dim)

    @staticmethod
    def backward(ctx, g):
        parts = [p.contiguous() for p in torch.chunk(g, ctx.n, dim=ctx.dim)]
        out = torch.empty_like(parts[0])
        dist.reduce_scatter(out, parts, group=ctx.group) if dist.get_backend(ctx.group) != "gloo" else _rs_gloo(out, parts, ctx.group)
        return out, None, None


def _rs_gloo(out, parts, group):
    """gloo has no reduce_scatter: all-reduce the stacked tensor and keep our slice."""
    full = torch.stack(parts)
    dist.all_reduce(full, group=group)
    out.copy_(full[dist.get_rank(group)])


class _RingGatherSeq(torch.autograd.Function):
    """Same result as ``_GatherSeq`` but the blocks hop around the ring (p2p mode)."""

    @staticmethod
    def forward(ctx, x, dim, group):
        n, r = dist.get_world_size(group), dist.get_rank(group)
        ctx.dim, ctx.group, ctx.n = dim, group, n
        ranks = dist.get_process_group_ranks(group)
        blocks = [None] * n
        blocks[r] = x.contiguous()
        cur = blocks[r]
        for hop in range(1, n):
            recv = torch.empty_like(cur)
            ops_ = [dist.P2POp(dist.isend, cur, ranks[(r + 1) % n], group), dist.P2POp(dist.irecv, recv, ranks[(r - 1) % n], group)]
            for w in dist.batch_isend_irecv(ops_):
                w.wait()
            blocks[(r - hop) % n] = recv
            cur = recv
        return torch.cat(blocks, dim=dim)

    backward = _GatherSeq.backward


class _SeqToHeads(torch.autograd.Function):
    """[b, h, s_local, d] -> [b, h/cp, s_local*cp, d] (rank-major sequence order) and back."""

    @staticmethod
    def forward(ctx, x, group, inverse):
        ctx.group, ctx.inverse = group, inverse
        return _a2a(x, group, inverse)

    @staticmethod
    def backward(ctx, g):
        return _a2a(g.contiguous(), ctx.group, not ctx.inverse), None, None


def _a2a(x, group, inverse):
    n = dist.get_world_size(group)
    b, h, s, d = x.shape
    if not inverse:      # scatter heads, gather sequence
        send = x.reshape(b, n, h // n, s, d).permute(1, 0, 2, 3, 4).contiguous()       # [n, b, h/n, s, d]
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send, group=group) if dist.get_backend(group) != "gloo" else _a2a_gloo(recv, send, group)
        return recv.permute(1, 2, 0, 3, 4).reshape(b, h // n, n * s, d)                # seq = rank-major
    send = x.reshape(b, h, n, s // n, d).permute(2, 0, 1, 3, 4).contiguous()            # [n, b, h, s/n, d]
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send, group=group) if dist.get_backend(group) != "gloo" else _a2a_gloo(recv, send, group)
    return recv.permute(1, 0, 2, 3, 4).reshape(b, n * h, s // n, d)


def _a2a_gloo(recv, send, group):
    n, r = dist.get_world_size(group), dist.get_rank(group)
    gathered = [torch.empty_like(send) for _ in range(n)]
    dist.all_gather(gathered, send, group=group)
    for src in range(n):
        recv[src].copy_(gathered[src][r])


# ------------------------------------------------------------------------------ attention
def cp_attention(q, k, v, scale, causal: bool, comm_type: str = "p2p", dropout_p: float = 0.0, training: bool = True,
                 attn_fn=None):
    """Attention over the full (cp-sharded) sequence.  q/k/v: local ``[b, h, s_local, d]``."""
    from .. import ops

    attn_fn = attn_fn or ops.attention
    group, cp, rank = _cp()
    if cp == 1:
        return attn_fn(q, k, v, scale, causal, dropout_p, training, None)
    s_local = q.shape[2]
    s_full = s_local * cp
    comm_type = (comm_type or "p2p").lower()
    heads_ok = q.shape[1] % cp == 0 and k.shape[1] % cp == 0
    hcp = ps.get_hierarchical_context_parallel_groups() if comm_type == "a2a+p2p" else None
    if hcp is not None:
        return _hierarchical_attention(q, k, v, scale, causal, dropout_p, training, attn_fn, hcp, cp, rank)
    if comm_type in ("a2a", "a2a+p2p") and heads_ok:
        order = gathered_order(s_full, cp, q.device)        # position of each gathered row
        to_nat = torch.argsort(order)                        # gathered row index for each natural position
        qf, kf, vf = (_SeqToHeads.apply(t.contiguous(), group, False).index_select(2, to_nat) for t in (q, k, v))
        o = attn_fn(qf, kf, vf, scale, causal, dropout_p, training, None)
        o = o.index_select(2, order)                         # natural -> gathered (rank-major) order
        return _SeqToHeads.apply(o.contiguous(), group, True)
    gather = _RingGatherSeq if comm_type == "p2p" else _GatherSeq
    kf = gather.apply(k.contiguous(), 2, group)
    vf = gather.apply(v.contiguous(), 2, group)
    if not causal:
        return attn_fn(q, kf, vf, scale, False, dropout_p, training, None)
    qpos = local_positions(s_full, cp, rank, q.device)
    kpos = gathered_order(s_full, cp, q.device)
    mask = (kpos.unsqueeze(0) > qpos.unsqueeze(1)).view(1, 1, s_local, s_full)   # True = masked out
    return attn_fn(q, kf, vf, scale, False, dropout_p, training, mask)


def _hierarchical_attention(q, k, v, scale, causal, dropout_p, training, attn_fn, hcp, cp, rank):
    """``a2a+p2p``: level 0 (inner, ``a`` ranks) exchanges sequence for heads -- every rank then holds ``h / a`` heads of the
    ``a * s_local`` tokens owned by its inner group --, level 1 (outer, ``p`` ranks) circulates those K/V blocks around a ring
    so each rank sees all keys for its heads; a position mask (from the zig-zag chunk ids) keeps causality; the inverse
    all-to-all restores ``[b, h, s_local, d]``."""
    inner, outer = hcp
    a = dist.get_world_size(inner) if not isinstance(inner, ps._SingleGroup) else 1
    p_ = dist.get_world_size(outer) if not isinstance(outer, ps._SingleGroup) else 1
    if q.shape[1] % a or k.shape[1] % a:
        raise ValueError(f"a2a+p2p context parallelism needs the head counts ({q.shape[1]}, {k.shape[1]}) divisible by the inner size {a}")
    s_local = q.shape[2]
    s_full = s_local * cp
    if a > 1:
        q, k, v = (_SeqToHeads.apply(t.contiguous(), inner, False) for t in (q, k, v))            # [b, h/a, a*s_local, d]
    if p_ > 1:
        kf = _RingGatherSeq.apply(k.contiguous(), 2, outer)
        vf = _RingGatherSeq.apply(v.contiguous(), 2, outer)
    else:
        kf, vf = k, v
    if causal:
        inner_idx, outer_idx = rank % a, rank // a                                                   # cp rank = outer * a + inner
        def group_pos(o):                                                                            # rows of inner group o, rank-major
            return torch.cat([local_positions(s_full, cp, o * a + i, q.device) for i in range(a)])
        qpos = group_pos(outer_idx)
        kpos = torch.cat([group_pos(o) for o in range(p_)])
        mask = (kpos.unsqueeze(0) > qpos.unsqueeze(1)).view(1, 1, qpos.numel(), kpos.numel())
        o = attn_fn(q, kf, vf, scale, False, dropout_p, training, mask)
    else:
        o = attn_fn(q, kf, vf, scale, False, dropout_p, training, None)
    return _SeqToHeads.apply(o.contiguous(), inner, True) if a > 1 else o
