"""MoE token dispatchers.

Parity: reference ``transformer/moe/token_dispatcher.py`` -- ``MoEAllGatherTokenDispatcher``
(:114-246), ``MoEAlltoAllTokenDispatcher`` (:248-706: permute -> all_to_all over EP -> optional
TP all-gather -> sort by local expert -> experts -> reverse) and the DeepEP-style ``flex``
dispatcher (:753-909) whose role here is taken by ``MoEPeerScatterDispatcher``: routed rows are
written straight into the expert owner's capacity-padded buffer by the sm_100a
``a2a_scatter_rows`` kernel over the EP symmetric arena, with device-side offsets (no
host-synchronised split sizes, reference ``_maybe_dtoh_and_synchronize :669``).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

from ..parallel import state as ps
from ..parallel.mappings import (_gather_along_first_dim, _reduce_scatter_along_first_dim, all_to_all,
                                 gather_from_sequence_parallel_region, reduce_scatter_to_sequence_parallel_region)
from .moe_utils import permute, sort_chunks_by_idxs, unpermute


def _ws(g):
    return 1 if (g is None or isinstance(g, ps._SingleGroup) or not dist.is_initialized()) else dist.get_world_size(g)


class MoETokenDispatcher:
    def __init__(self, config):
        self.config = config
        self.ep_group = ps.get_expert_model_parallel_group(check_initialized=False)
        self.ep_size = ps.get_expert_model_parallel_world_size()
        self.ep_rank = ps.get_expert_model_parallel_rank()
        self.etp_group = ps.get_expert_tensor_parallel_group(check_initialized=False)
        self.etp_size = ps.get_expert_tensor_parallel_world_size()
        self.shared_experts = None

    def set_shared_experts(self, shared):
        self.shared_experts = shared


class MoEAllGatherTokenDispatcher(MoETokenDispatcher):
    """Gather every token (and its routing) across the TPxEP group; each rank keeps the tokens
    routed to its local experts; outputs are reduce-scattered back."""

    def __init__(self, num_local_experts, local_expert_indices, config):
        super().__init__(config)
        self.num_local_experts, self.local_expert_indices = num_local_experts, local_expert_indices
        self.tp_ep_group = ps.get_expert_tensor_and_model_parallel_group(check_initialized=False)
        self.tp_ep_size = _ws(self.tp_ep_group)

    def token_permutation(self, hidden_states, probs, routing_map):
        self.hidden_shape = hidden_states.shape
        x = hidden_states.view(-1, self.hidden_shape[-1])
        if self.tp_ep_size > 1:
            routing_map = _gather_along_first_dim(routing_map.to(torch.uint8), self.tp_ep_group).bool()
            probs = _GatherFirstDimAutograd.apply(probs, self.tp_ep_group)
            x = _GatherFirstDimAutograd.apply(x, self.tp_ep_group)
        self.hidden_shape_before_permute = x.shape
        lo, hi = self.local_expert_indices[0], self.local_expert_indices[-1] + 1
        self.local_map = routing_map[:, lo:hi].contiguous()
        self.local_probs = probs[:, lo:hi].contiguous()
        tokens_per_expert = self.local_map.sum(dim=0).long().cpu()
        permuted, self.rev = permute(x, self.local_map)
        return permuted, tokens_per_expert

    def token_unpermutation(self, hidden_states, bias=None):
        out = unpermute(hidden_states, self.rev, self.hidden_shape_before_permute, self.local_probs, self.local_map)
        if self.tp_ep_size > 1:
            out = _ReduceScatterFirstDimAutograd.apply(out, self.tp_ep_group)
        return out.view(self.hidden_shape), None


class _GatherFirstDimAutograd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        return _gather_along_first_dim(x, group)

    @staticmethod
    def backward(ctx, g):
        return _reduce_scatter_along_first_dim(g.contiguous(), ctx.group), None


class _ReduceScatterFirstDimAutograd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        return _reduce_scatter_along_first_dim(x.contiguous(), group)

    @staticmethod
    def backward(ctx, g):
        return _gather_along_first_dim(g.contiguous(), ctx.group), None


class MoEAlltoAllTokenDispatcher(MoETokenDispatcher):
    def __init__(self, num_local_experts, local_expert_indices, config):
        super().__init__(config)
        self.num_local_experts, self.local_expert_indices = num_local_experts, local_expert_indices
        self.num_experts = config.num_moe_experts
        # chunk (src rank, local expert) -> (local expert, src rank) ordering tables
        n = self.num_experts * self.etp_size
        self.sort_input_by_local_experts = torch.arange(n).reshape(-1, num_local_experts).T.ravel().tolist()
        self.restore_output_by_local_experts = torch.arange(n).reshape(num_local_experts, -1).T.ravel().tolist()

    def _meta(self, routing_map):
        E, ep, nl = self.num_experts, self.ep_size, self.num_local_experts
        local_counts = routing_map.sum(dim=0).long()                      # [E] tokens I send per global expert
        self.input_splits = local_counts.view(ep, nl).sum(dim=1).cpu().tolist() if ep > 1 else None
        if ep > 1 or self.etp_size > 1:
            gathered = local_counts.view(1, E)
            if ep > 1:
                gathered = _gather_along_first_dim(local_counts.view(1, E), self.ep_group)   # [ep, E]
            mine = gathered[:, self.local_expert_indices[0]:self.local_expert_indices[-1] + 1]  # [ep, nl]
            if self.etp_size > 1:
                mine_all = _gather_along_first_dim(mine.contiguous().view(1, -1), self.etp_group).view(self.etp_size, -1, nl)
                self.output_splits_tp = mine_all.sum(dim=(1, 2)).cpu().tolist()
                self.num_global_tokens_per_local_expert = mine_all.reshape(-1, nl)  # [(etp*ep), nl]
            else:
                self.output_splits_tp = None
                self.num_global_tokens_per_local_expert = mine
            self.output_splits = mine.sum(dim=1).cpu().tolist() if ep > 1 else None
            tokens_per_expert = self.num_global_tokens_per_local_expert.sum(dim=0).cpu()
            self.chunk_sizes = self.num_global_tokens_per_local_expert.reshape(-1).cpu()
        else:
            self.output_splits = self.output_splits_tp = None
            self.num_global_tokens_per_local_expert = None
            tokens_per_expert = local_counts.cpu()
        return tokens_per_expert

    def token_permutation(self, hidden_states, probs, routing_map):
        self.hidden_shape = hidden_states.shape
        self.probs, self.routing_map = probs, routing_map
        x = hidden_states.view(-1, self.hidden_shape[-1])
        tokens_per_expert = self._meta(routing_map)
        self.hidden_shape_before_permute = x.shape
        permuted, self.rev = permute(x, routing_map)
        if self.ep_size > 1:
            permuted = all_to_all(self.ep_group, permuted, self.output_splits, self.input_splits)
        if self.etp_size > 1:
            permuted = _GatherFirstDimVarAutograd.apply(permuted, self.etp_group, self.output_splits_tp)
        if self.num_local_experts > 1 and self.num_global_tokens_per_local_expert is not None:
            n_chunks = self.chunk_sizes.numel()
            order = torch.arange(n_chunks).reshape(-1, self.num_local_experts).T.ravel().tolist()
            self._restore = torch.arange(n_chunks).reshape(self.num_local_experts, -1).T.ravel().tolist()
            self._sorted_sizes = self.chunk_sizes[order]
            permuted = sort_chunks_by_idxs(permuted, self.chunk_sizes, order)
        return permuted, tokens_per_expert

    def token_unpermutation(self, hidden_states, bias=None):
        x = hidden_states
        if self.num_local_experts > 1 and self.num_global_tokens_per_local_expert is not None:
            x = sort_chunks_by_idxs(x, self._sorted_sizes, self._restore)
        if self.etp_size > 1:
            x = _ReduceScatterFirstDimVarAutograd.apply(x, self.etp_group, self.output_splits_tp)
        if self.ep_size > 1:
            x = all_to_all(self.ep_group, x, self.input_splits, self.output_splits)
        out = unpermute(x, self.rev, self.hidden_shape_before_permute, self.probs, self.routing_map)
        return out.view(self.hidden_shape), None


class _GatherFirstDimVarAutograd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, group, sizes):
        ctx.group, ctx.sizes = group, sizes
        return _gather_along_first_dim(x, group, sizes)

    @staticmethod
    def backward(ctx, g):
        return _reduce_scatter_along_first_dim(g.contiguous(), ctx.group, ctx.sizes), None, None


class _ReduceScatterFirstDimVarAutograd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, group, sizes):
        ctx.group, ctx.sizes = group, sizes
        return _reduce_scatter_along_first_dim(x.contiguous(), group, sizes)

    @staticmethod
    def backward(ctx, g):
        return _gather_along_first_dim(g.contiguous(), ctx.group, ctx.sizes), None, None


class RaggedTokens:
    """Expert-side token buffer of the fused dispatcher: ``data`` [capacity, h] in the 256-row padded per-expert layout
    described by ``layout`` (``ops/moe.py``); ``probs`` is the router's dense probability tensor (carries the gradient).
    ``combined`` is set by experts that already applied the probabilities and summed the copies of every token
    (``GroupedMLP`` does both inside the FC2 kernel + gather-sum)."""

    def __init__(self, data, layout, probs, combined=None):
        self.data, self.layout, self.probs, self.combined = data, layout, probs, combined

    @property
    def shape(self):
        return self.data.shape


class _Combine(torch.autograd.Function):
    """out[t] = sum_e p[t, e] * y[row(t, e)] for experts that return un-scaled ragged rows (SequentialMLP)."""

    @staticmethod
    def forward(ctx, y, probs, lay):
        from ..ops import moe as mops

        ctx.lay = lay
        ctx.save_for_backward(y)
        ctx.probs_shape, ctx.probs_dtype = probs.shape, probs.dtype
        z = (y.float() * lay.row_prob.unsqueeze(-1)).to(y.dtype)
        return mops._gather_sum(z, lay)

    @staticmethod
    def backward(ctx, dout):
        from ..ops import moe as mops

        (y,) = ctx.saved_tensors
        lay = ctx.lay
        dprobs = torch.zeros(ctx.probs_shape, dtype=torch.float32, device=dout.device)
        g = mops._gather_rows(dout.contiguous(), lay, lay.E_total)                # dOut re-permuted, un-scaled
        valid = (lay.row_slot >= 0)
        d = (g.float() * y.float()).sum(-1)
        dprobs.view(-1)[lay.row_slot[valid].long()] = d[valid]
        dy = (g.float() * lay.row_prob.unsqueeze(-1)).to(y.dtype)
        return dy, dprobs.to(ctx.probs_dtype), None


class MoEFusedDispatcher(MoETokenDispatcher):
    """``--moe-token-dispatcher-type fused``: device-side routing layout + grouped tcgen05 expert GEMMs, zero host syncs.

    ``token_permutation`` builds the ragged layout on the device (``moe.cu``: stable per-expert positions, 256-row padded
    segments, row <-> (token, expert) maps) and gathers the token rows into it; the experts consume it as ONE grouped GEMM
    per matrix; ``token_unpermutation`` is a gather-sum.  With expert tensor parallelism the tokens / routing of the
    TP group are all-gathered in front (every TP rank holds 1/etp of every expert's FFN) and the partial outputs are
    reduce-scattered behind, as in the reference's dispatcher (token_dispatcher.py:560-640).  Expert parallelism > 1 goes
    through the EP symmetric arena (``MoEArenaDispatcher``)."""

    def __init__(self, num_local_experts, local_expert_indices, config):
        super().__init__(config)
        self.num_local_experts, self.local_expert_indices = num_local_experts, local_expert_indices
        self.num_experts = config.num_moe_experts
        if self.ep_size > 1:
            raise ValueError("MoEFusedDispatcher handles expert-parallel size 1; use MoEArenaDispatcher (flex) for ep > 1")

    def token_permutation(self, hidden_states, probs, routing_map):
        from ..ops import moe as mops

        self.hidden_shape = hidden_states.shape
        x = hidden_states.reshape(-1, self.hidden_shape[-1])
        if self.etp_size > 1:
            routing_map = _gather_along_first_dim(routing_map.to(torch.uint8), self.etp_group).bool()
            probs = _GatherFirstDimAutograd.apply(probs.contiguous(), self.etp_group)
            x = _GatherFirstDimAutograd.apply(x.contiguous(), self.etp_group)
        lay = mops.build_layout(routing_map.contiguous(), probs, self.local_expert_indices[0], self.num_local_experts,
                                self.config.moe_router_topk)
        return RaggedTokens(mops.permute(x, lay), lay, probs), lay.counts

    def token_unpermutation(self, expert_out, bias=None):
        if isinstance(expert_out, RaggedTokens):
            out = expert_out.combined if expert_out.combined is not None else \
                _Combine.apply(expert_out.data, expert_out.probs, expert_out.layout)
        else:
            raise TypeError("MoEFusedDispatcher expects the experts to return RaggedTokens")
        if self.etp_size > 1:
            out = _ReduceScatterFirstDimAutograd.apply(out, self.etp_group)
        return out.view(self.hidden_shape), None


class _EPCall:
    """State of one dispatch -> experts -> combine round trip over the EP arena (one MoE layer call)."""

    def __init__(self, disp, send_lay, lay, row_idx, n_tokens, hidden):
        self.disp, self.send_lay, self.lay, self.row_idx = disp, send_lay, lay, row_idx
        self.n_tokens, self.hidden = n_tokens, hidden
        self.dprob_row = None        # filled by the FC2 backward, shipped home by the dispatch backward


class _EPDispatch(torch.autograd.Function):
    """tokens [n, h] -> rows of my local experts [cap, h] (256-row padded ragged layout) over NVLink.
    backward: d(rows) travel home the same way they will in the combine (push + gather-sum), d(probs) ride along."""

    @staticmethod
    def forward(ctx, x, probs, disp, send_lay):
        call = disp._dispatch_forward(x, probs, send_lay)
        ctx.call = call
        ctx.probs_dtype = probs.dtype
        disp._last_call = call
        return call.xp

    @staticmethod
    def backward(ctx, dxp):
        call = ctx.call
        disp = call.disp
        ret_x, ret_f = disp._push_home(dxp.contiguous(), call, call.dprob_row)
        from ..ops import ext

        dx = ext().moe_gather_sum(ret_x, call.send_lay.dest_row)
        dprobs = ext().moe_scatter_dprobs(ret_f, call.send_lay.dest_row) if call.dprob_row is not None else None
        return dx, (dprobs.to(ctx.probs_dtype) if dprobs is not None else None), None, None


class _GroupedLinearCombineEP(torch.autograd.Function):
    """Expert FC2 on the owner (router probability in the GEMM epilogue) + combine over NVLink + gather-sum at home."""

    @staticmethod
    def forward(ctx, h, weight, call, stride, fuse_wgrad):
        from ..ops import ext
        from ..ops import moe as mops

        ctx.call, ctx.stride, ctx.fuse = call, stride, fuse_wgrad
        z = mops.grouped_gemm(h, weight, call.lay, stride, row_scale=call.lay.row_prob)
        ret_x, _ = call.disp._push_home(z, call, None)
        out = ext().moe_gather_sum(ret_x, call.send_lay.dest_row)
        ctx.save_for_backward(h, weight, z)
        return out

    @staticmethod
    def backward(ctx, dout):
        from ..ops import ext
        from ..ops import moe as mops

        h, weight, z = ctx.saved_tensors
        call = ctx.call
        disp = call.disp
        stage = disp._dispatch_rows_only(dout.contiguous(), call)          # dOut rows land in my staging again
        dprob_stage = torch.zeros(disp.ep_size * disp.Rs, dtype=torch.float32, device=dout.device)
        dy = ext().moe_gather_rows(stage, call.lay.row_slot, 1, call.lay.row_prob, call.lay.offsets, call.lay.E_local, z,
                                   dprob_stage)
        call.dprob_row = dprob_stage[call.lay.row_slot.clamp(min=0).long()]
        dh = mops.grouped_gemm(dy, weight, call.lay, ctx.stride, b_mn=True) if ctx.needs_input_grad[0] else None
        dw = mops.grouped_wgrad(dy, h, weight, call.lay, ctx.stride, ctx.fuse) if ctx.needs_input_grad[1] else None
        return dh, dw, None, None, None


class MoEArenaDispatcher(MoETokenDispatcher):
    """``--moe-token-dispatcher-type flex`` with expert parallelism: sync-free dispatch / combine over the EP symmetric arena.

    Every source rank writes its routed rows straight into the expert owner's staging buffer with NVLink stores
    (``moe.cu: ep_dispatch_kernel``), together with its per-expert counts, the slot each row returns to and the router
    probability; the owner turns the counts into the ragged 256-row padded layout on the device (``ep_recv_layout``),
    runs the grouped tcgen05 GEMMs over it and pushes the (probability-scaled) outputs back to the slots the sources
    named (``ep_combine_push``); the source finishes with a gather-sum.  No host synchronisation, no NCCL, no padding to a
    capacity factor: the only static bound is the compact buffer size (``B200_MOE_SLACK`` x the balanced load; an overflow
    raises a device flag that ``check_overflow()`` reports).  Parity: reference token_dispatcher.py:248-706 (alltoall) and
    the DeepEP ``flex`` dispatcher fused_a2a.py:66-202."""

    def __init__(self, num_local_experts, local_expert_indices, config):
        import os

        super().__init__(config)
        self.num_local_experts, self.local_expert_indices = num_local_experts, local_expert_indices
        self.num_experts = config.num_moe_experts
        self.slack = float(os.environ.get("B200_MOE_SLACK", "2.0"))
        self._ready = False
        self._last_call = None

    # ---- arena ------------------------------------------------------------------------------------------------
    def _setup(self, n, hidden, dtype, device):
        from ..comm.arena import get_arena

        ep, El, k = self.ep_size, self.num_local_experts, self.config.moe_router_topk
        self.n, self.hidden = n, hidden
        self.Rs = (n * min(k, El) + 7) // 8 * 8                              # staging rows per source
        self.cap_send = n * min(k, self.num_experts)
        bal = n * k                                                          # balanced load: what I send == what I receive
        self.cap = (int(self.slack * bal) + 255) // 256 * 256 + 256 * El
        sizes = {"ep_stage_x": ep * self.Rs * hidden * 2, "ep_stage_idx": ep * self.Rs * 4, "ep_stage_prob": ep * self.Rs * 4,
                 "ep_recv_counts": ep * El * 4 + 64, "ep_flags_d": 256, "ep_flags_c": 256,
                 "ep_ret_x": self.cap_send * hidden * 2, "ep_ret_f": self.cap_send * 4}
        total = sum((v + 1023) // 1024 * 1024 for v in sizes.values()) + (1 << 20)
        self.arena = get_arena("ep", self.ep_group, total)
        self.reg = {name: self.arena.region(name, nbytes) for name, nbytes in sizes.items()}
        a = self.arena
        self.stage_x = a.local(self.reg["ep_stage_x"], (ep * self.Rs, hidden), dtype)
        self.stage_idx = a.local(self.reg["ep_stage_idx"], (ep * self.Rs,), torch.int32)
        self.stage_prob = a.local(self.reg["ep_stage_prob"], (ep * self.Rs,), torch.float32)
        self.ret_x = a.local(self.reg["ep_ret_x"], (self.cap_send, hidden), dtype)
        self.ret_f = a.local(self.reg["ep_ret_f"], (self.cap_send,), torch.float32)
        self.done = torch.zeros(2, dtype=torch.int32, device=device)
        self.overflow = torch.zeros(1, dtype=torch.int32, device=device)
        self._ready = True

    def check_overflow(self) -> bool:
        """True when some call received more rows than the compact buffer holds (host sync; call it at log intervals)."""
        return self._ready and bool(self.overflow.item())

    # ---- wire protocol ------------------------------------------------------------------------------------------
    def _send(self, x, send_lay, probs_f32):
        from ..ops import ext

        a, r = self.arena, self.reg
        epoch = a.next_epoch(r["ep_flags_d"])
        ext().moe_ep_dispatch(x, send_lay.offsets, send_lay.dest_row, send_lay.counts, probs_f32, a.ptrs(r["ep_stage_x"]),
                              a.ptrs(r["ep_stage_idx"]), a.ptrs(r["ep_stage_prob"]), a.ptrs(r["ep_recv_counts"]),
                              a.ptrs(r["ep_flags_d"]), self.num_local_experts, a.rank, self.Rs, epoch, self.done[0:1])
        return epoch

    def _dispatch_forward(self, x, probs, send_lay):
        from ..ops import ext
        from ..ops.moe import RaggedLayout

        if not self._ready:
            self._setup(x.shape[0], x.shape[1], x.dtype, x.device)
        if x.shape[0] != self.n or x.shape[1] != self.hidden:
            raise RuntimeError("MoEArenaDispatcher: token count / hidden size changed after the arena was sized")
        a, r = self.arena, self.reg
        epoch = self._send(x.contiguous(), send_lay, probs.detach().float().contiguous())
        El = self.num_local_experts
        offsets, counts, row_stage = ext().moe_ep_recv_layout(a.local_ptr + r["ep_flags_d"].offset,
                                                              a.local_ptr + r["ep_recv_counts"].offset, self.ep_size, El, self.Rs,
                                                              256, self.cap, epoch, self.overflow)
        safe = row_stage.clamp(min=0).long()
        valid = row_stage >= 0
        row_prob = self.stage_prob[safe] * valid
        row_idx = self.stage_idx[safe]
        lay = RaggedLayout(x.shape[0], 1, 0, El, self.cap, counts, offsets, row_stage, row_prob, None)
        call = _EPCall(self, send_lay, lay, row_idx, x.shape[0], x.shape[1])
        call.xp = ext().moe_gather_rows(self.stage_x, row_stage, 1, None, offsets, El, None, None)
        return call

    def _dispatch_rows_only(self, x, call):
        """Backward of the combine: the same routing again, payload = dOut rows; returns my staging buffer."""
        from ..ops import ext

        a, r = self.arena, self.reg
        epoch = self._send(x, call.send_lay, None)
        ext().moe_ep_wait(a.local_ptr + r["ep_flags_d"].offset, self.ep_size, epoch)
        return self.stage_x

    def _push_home(self, z, call, side):
        from ..ops import ext

        a, r = self.arena, self.reg
        epoch = a.next_epoch(r["ep_flags_c"])
        ext().moe_ep_combine_push(z, call.lay.row_slot, call.row_idx, side, call.lay.offsets, call.lay.E_local, self.Rs,
                                  a.ptrs(r["ep_ret_x"]), a.ptrs(r["ep_ret_f"]), a.ptrs(r["ep_flags_c"]), a.rank, epoch,
                                  self.done[1:2])
        ext().moe_ep_wait(a.local_ptr + r["ep_flags_c"].offset, self.ep_size, epoch)
        return self.ret_x, self.ret_f

    # ---- dispatcher API -------------------------------------------------------------------------------------------
    def token_permutation(self, hidden_states, probs, routing_map):
        from ..ops import moe as mops

        if not hidden_states.is_cuda:
            raise NotImplementedError("the arena (flex) dispatcher needs CUDA peers; use --moe-token-dispatcher-type alltoall on CPU")
        self.hidden_shape = hidden_states.shape
        x = hidden_states.reshape(-1, self.hidden_shape[-1])
        if self.etp_size > 1:
            routing_map = _gather_along_first_dim(routing_map.to(torch.uint8), self.etp_group).bool()
            probs = _GatherFirstDimAutograd.apply(probs.contiguous(), self.etp_group)
            x = _GatherFirstDimAutograd.apply(x.contiguous(), self.etp_group)
        send_lay = mops.build_layout(routing_map.contiguous(), probs, 0, self.num_experts, self.config.moe_router_topk, pad=1)
        xp = _EPDispatch.apply(x, probs, self, send_lay)
        call, self._last_call = self._last_call, None     # do not keep the call (received tokens, layouts) alive across steps
        rt = RaggedTokens(xp, call.lay, probs)
        rt.ep_call = call
        return rt, call.lay.counts

    def token_unpermutation(self, expert_out, bias=None):
        if not (isinstance(expert_out, RaggedTokens) and expert_out.combined is not None):
            raise TypeError("MoEArenaDispatcher needs GroupedMLP experts (--moe-grouped-gemm)")
        out = expert_out.combined
        if self.etp_size > 1:
            out = _ReduceScatterFirstDimAutograd.apply(out, self.etp_group)
        return out.view(self.hidden_shape), None


class _PeerScatter(torch.autograd.Function):
    """rows -> owner buffers via the fused NVLink scatter kernel; backward = the reverse scatter."""

    @staticmethod
    def forward(ctx, x, disp, direction):
        ctx.disp, ctx.direction = disp, direction
        return disp._scatter(x, direction)

    @staticmethod
    def backward(ctx, g):
        return ctx.disp._scatter(g.contiguous(), 1 - ctx.direction), None, None


class MoEPeerScatterDispatcher(MoETokenDispatcher):
    """``--moe-token-dispatcher-type flex``: sync-free dispatch into capacity-padded expert buffers.

    Every rank computes, on the device, the destination (owner rank, row) of each routed copy from
    an all-gather of the per-expert counts (tiny, stays on the GPU); rows beyond an expert's
    capacity are dropped (``dst_rank = -1``).  The expert sees a static ``[E_local * ep * cap, h]``
    buffer (zero rows where nothing landed), so expert GEMM shapes never depend on the routing and
    nothing is copied to the host."""

    def __init__(self, num_local_experts, local_expert_indices, config):
        super().__init__(config)
        self.num_local_experts, self.local_expert_indices = num_local_experts, local_expert_indices
        self.num_experts = config.num_moe_experts
        self.capacity_factor = config.moe_expert_capacity_factor or 1.25
        self._arena = None
        self._epoch = 0

    def _get_arena(self, nbytes):
        from ..comm.arena import get_arena

        a = get_arena("ep", self.ep_group, nbytes)
        return a

    def _scatter(self, x, direction):
        """direction 0: tokens -> expert buffers; 1: expert buffers -> token-copy rows."""
        from .. import ops

        h = x.shape[1]
        rows_out = self.buf_rows if direction == 0 else self.n_copies
        if self.ep_size > 1 and not x.is_cuda:
            raise NotImplementedError("the peer-scatter (flex) dispatcher needs CUDA peers for ep > 1; "
                                      "use --moe-token-dispatcher-type alltoall on CPU")
        if self.ep_size == 1:
            out = torch.zeros(rows_out, h, dtype=x.dtype, device=x.device)
            keep = self.dst_rank >= 0
            rows = self.dst_row[keep].long()
            if direction == 0:
                out[rows] = x[keep]           # copy i -> buffer row dst_row[i]
            else:
                out[keep] = x[rows]           # buffer row -> copy i
            return out
        arena = self._get_arena(4 * max(self.buf_rows, self.n_copies) * h * 2 + (1 << 20))
        slot = self._epoch & 1
        reg = arena.region(f"moe_buf{slot}_{direction}", max(self.buf_rows, self.n_copies) * h * 2)
        flags = arena.region(f"moe_flags{slot}_{direction}", 64 * 8 * 4)
        ep = arena.next_epoch(flags)
        self._epoch += 1
        local = arena.local(reg, (rows_out, h), x.dtype)
        local.zero_()
        arena.barrier()  # everybody's buffer is cleared before rows start landing
        if direction == 0:
            ops.ext().a2a_scatter_rows(x.contiguous(), self.dst_rank, self.dst_row, arena.ptrs(reg), arena.ptrs(flags),
                                       arena.rank, arena.size, ep, 32)
        else:
            ops.ext().a2a_scatter_rows(x.contiguous(), self.back_rank, self.back_row, arena.ptrs(reg), arena.ptrs(flags),
                                       arena.rank, arena.size, ep, 32)
        return local.clone()

    def token_permutation(self, hidden_states, probs, routing_map):
        self.hidden_shape = hidden_states.shape
        self.probs, self.routing_map = probs, routing_map
        x = hidden_states.view(-1, self.hidden_shape[-1])
        n, E, nl, ep = x.shape[0], self.num_experts, self.num_local_experts, self.ep_size
        dev = x.device
        cap = int(-(-n * self.config.moe_router_topk * self.capacity_factor // E))
        cap = (cap + 127) // 128 * 128
        self.cap = cap
        rm = routing_map.bool().T.contiguous()                               # [E, n]
        tok = torch.arange(n, device=dev).unsqueeze(0).expand(E, -1)
        self.copy_idx = tok.masked_select(rm).int()                          # source token of each routed copy
        exp_of_copy = torch.arange(E, device=dev).unsqueeze(1).expand(-1, n).masked_select(rm)
        counts = rm.sum(dim=1).int()                                         # [E]
        pos = (torch.cumsum(rm.int(), dim=1) - 1).masked_select(rm).int()    # position inside (me, expert)
        owner = (exp_of_copy // nl).int()
        le = (exp_of_copy % nl).int()
        # buffer layout on the owner: [local expert][src rank][cap]
        self.dst_row = ((le * ep + self.ep_rank) * cap + pos).int()
        self.dst_rank = torch.where(pos < cap, owner, torch.full_like(owner, -1)).int()
        self.n_copies = int(self.copy_idx.numel())
        self.buf_rows = nl * ep * cap
        # reverse map: for every buffer row, which (rank, copy row) it came from (-1 = empty)
        if ep > 1 and x.is_cuda:
            back_rank = torch.full((self.buf_rows,), -1, dtype=torch.int32, device=dev)
            back_row = torch.zeros(self.buf_rows, dtype=torch.int32, device=dev)
            meta = torch.stack([self.dst_rank, self.dst_row, torch.arange(self.n_copies, device=dev, dtype=torch.int32)])
            gathered = [torch.empty_like(meta) for _ in range(ep)] if False else None
            sizes = torch.tensor([self.n_copies], device=dev)
            all_sizes = [torch.zeros_like(sizes) for _ in range(ep)]
            dist.all_gather(all_sizes, sizes, group=self.ep_group)
            mx = int(max(int(s) for s in all_sizes))
            pad = torch.full((3, mx), -1, dtype=torch.int32, device=dev)
            pad[:, :self.n_copies] = meta
            allm = [torch.empty_like(pad) for _ in range(ep)]
            dist.all_gather(allm, pad, group=self.ep_group)
            for r, m in enumerate(allm):
                sel = m[0] == self.ep_rank
                back_rank[m[1][sel].long()] = r
                back_row[m[1][sel].long()] = m[2][sel]
            self.back_rank, self.back_row = back_rank, back_row
        copies = x.index_select(0, self.copy_idx.long())
        buf = _PeerScatter.apply(copies, self, 0)
        tokens_per_expert = torch.full((nl,), ep * cap, dtype=torch.long)
        return buf, tokens_per_expert

    def token_unpermutation(self, hidden_states, bias=None):
        copies = _PeerScatter.apply(hidden_states, self, 1)                   # [n_copies, h]
        n, h = self.hidden_shape[0] * self.hidden_shape[1] if len(self.hidden_shape) == 3 else self.hidden_shape[0], self.hidden_shape[-1]
        p = self.probs.T.contiguous().masked_select(self.routing_map.bool().T.contiguous())
        out = torch.zeros(n, h, dtype=copies.dtype, device=copies.device)
        out.index_add_(0, self.copy_idx.long(), copies * p.unsqueeze(-1).to(copies.dtype))
        return out.view(self.hidden_shape), None
