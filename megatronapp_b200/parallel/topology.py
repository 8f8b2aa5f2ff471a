"""Pure rank arithmetic for the process topology.

A :class:`Topology` is an immutable description of how ``world_size`` ranks are
factored along the parallel axes (tp, cp, ep, dp, pp and the MegaFBD ``fb``
pairing).  Everything here is plain integer math -- no ``torch.distributed`` --
so the layout logic is testable on a laptop.

Capability parity: reference ``megatron/core/parallel_state.py:1170-1246``
(``RankGenerator``, order string ``tp-cp-ep-dp-pp``), ``:1272-2056`` (group
layout), ``:444-512`` (MegaFBD forward/backward pairing and half-DP groups).
The design is different: one mixed-radix coordinate system, groups derived by
"vary these axes, fix the others".
"""
from __future__ import annotations

from dataclasses import dataclass, field
from functools import lru_cache
from typing import Dict, List, Sequence, Tuple

AXES = ("tp", "cp", "ep", "dp", "pp")


def _parse_order(order: str) -> Tuple[str, ...]:
    names = tuple(tok for tok in order.split("-") if tok)
    for n in names:
        if n not in AXES:
            raise ValueError(f"unknown axis {n!r} in order {order!r}")
    # axes not mentioned are appended (size must be 1 for them)
    return names + tuple(a for a in AXES if a not in names)


@dataclass(frozen=True)
class AxisLayout:
    """Mixed-radix layout: first axis in ``order`` varies fastest."""

    sizes: Tuple[Tuple[str, int], ...]  # ordered (name, size)
    offset: int = 0

    @property
    def world(self) -> int:
        w = 1
        for _, s in self.sizes:
            w *= s
        return w

    def stride(self, axis: str) -> int:
        st = 1
        for n, s in self.sizes:
            if n == axis:
                return st
            st *= s
        raise KeyError(axis)

    def size(self, axis: str) -> int:
        return dict(self.sizes)[axis]

    def coords(self, rank: int) -> Dict[str, int]:
        r = rank - self.offset
        out = {}
        for n, s in self.sizes:
            out[n] = r % s
            r //= s
        return out

    def rank_of(self, coords: Dict[str, int]) -> int:
        r, st = 0, 1
        for n, s in self.sizes:
            r += coords.get(n, 0) * st
            st *= s
        return r + self.offset

    def groups(self, token: str) -> List[List[int]]:
        """All groups obtained by varying the axes in ``token`` ("tp-pp" ...)
        and fixing the rest.  Inside a group, ranks are ordered with the
        earliest-in-order axis fastest (same convention as the reference)."""
        vary = [t for t in token.split("-") if t]
        names = [n for n, _ in self.sizes]
        for v in vary:
            if v not in names:
                raise KeyError(v)
        fixed = [n for n in names if n not in vary]
        vary_sorted = [n for n in names if n in vary]
        out: List[List[int]] = []
        fixed_sizes = [self.size(n) for n in fixed]
        nfixed = 1
        for s in fixed_sizes:
            nfixed *= s
        for fi in range(nfixed):
            c = {}
            t = fi
            for n, s in zip(fixed, fixed_sizes):
                c[n] = t % s
                t //= s
            grp = []
            nvary = 1
            for n in vary_sorted:
                nvary *= self.size(n)
            for vi in range(nvary):
                t = vi
                cc = dict(c)
                for n in vary_sorted:
                    s = self.size(n)
                    cc[n] = t % s
                    t //= s
                grp.append(self.rank_of(cc))
            out.append(grp)
        return out


@dataclass(frozen=True)
class Topology:
    """Immutable parallel topology.

    ``dp`` is the *full* data-parallel size.  With ``fbd=True`` (MegaFBD) the
    DP axis is split into pairs: even replica index = forward rank, odd =
    backward rank; the *effective* DP for gradient averaging is ``dp // 2``.
    """

    world_size: int
    tp: int = 1
    pp: int = 1
    cp: int = 1
    ep: int = 1
    etp: int | None = None  # expert tensor parallel (defaults to tp)
    vpp: int | None = None
    order: str = "tp-cp-ep-dp-pp"
    fbd: bool = False
    # MegaFBD with the forward tensor parallelism COLLAPSED (reference ``--ignore-forward-tensor-parallel``,
    # parallel_state.py:896-964): every (pipeline stage, replica) is a block of 1 + tp ranks -- ONE forward rank that
    # holds the un-sharded stage and runs true TP=1 math, followed by the tp backward ranks of that stage.
    fbd_collapse: bool = False
    num_distributed_optimizer_instances: int = 1
    split_rank: int | None = None   # first pipeline stage of the decoder (encoder-decoder models)
    # hierarchical context parallelism (``--hierarchical-context-parallel-sizes a p``, cp = a * p): level 0 = inner groups of
    # ``a`` consecutive cp ranks (all-to-all), level 1 = outer groups of ``p`` ranks with the same inner index (p2p ring)
    hcp_sizes: tuple | None = None

    def __post_init__(self):
        if self.split_rank is not None and not (0 < self.split_rank < self.pp):
            raise ValueError(f"pipeline split rank {self.split_rank} must lie strictly inside [1, pp={self.pp})")
        if self.fbd_collapse:
            if not self.fbd:
                raise ValueError("fbd_collapse needs fbd=True")
            if self.cp != 1 or self.ep != 1 or (self.etp not in (None, self.tp)):
                raise ValueError("collapsed MegaFBD does not support cp / ep > 1 or a separate expert TP size")
            blk = (self.tp + 1) * self.pp
            if self.world_size % blk != 0:
                raise ValueError(f"world_size ({self.world_size}) must be a multiple of (tp + 1) * pp = {blk}")
            object.__setattr__(self, "dp", self.world_size // blk)      # replicas; every replica has its own forward rank
            object.__setattr__(self, "etp", self.tp)
            object.__setattr__(self, "edp", self.world_size // blk)
            return
        mp = self.tp * self.pp * self.cp
        if self.world_size % mp != 0:
            raise ValueError(
                f"world_size ({self.world_size}) is not divisible by tp*pp*cp ({mp})"
            )
        dp = self.world_size // mp
        object.__setattr__(self, "dp", dp)
        etp = self.tp if self.etp is None else self.etp
        object.__setattr__(self, "etp", etp)
        emp = etp * self.ep * self.pp
        if self.world_size % emp != 0:
            raise ValueError(
                f"world_size ({self.world_size}) is not divisible by etp*ep*pp ({emp})"
            )
        object.__setattr__(self, "edp", self.world_size // emp)
        if self.ep > 1 and (dp * self.cp * self.tp) % (self.ep * etp) != 0:
            raise ValueError("ep*etp must divide dp*cp*tp")
        if self.vpp is not None and self.vpp > 1 and self.pp < 2:
            raise ValueError("virtual pipeline needs pp >= 2")
        if self.hcp_sizes:
            prod = 1
            for x in self.hcp_sizes:
                prod *= int(x)
            if prod != self.cp or len(self.hcp_sizes) != 2:
                raise ValueError(f"hierarchical context parallel sizes {self.hcp_sizes} must be two levels whose product is cp={self.cp}")
        if self.fbd:
            if dp % 2 != 0:
                raise ValueError("MegaFBD needs an even data-parallel size")
            if self.cp != 1 or self.ep != 1:
                raise ValueError("MegaFBD does not support cp/ep > 1")
        if dp % self.num_distributed_optimizer_instances != 0:
            raise ValueError("dp must be divisible by num_distributed_optimizer_instances")

    # ---- layouts -------------------------------------------------------
    @property
    def dense(self) -> AxisLayout:
        names = _parse_order(self.order)
        sz = {"tp": self.tp, "cp": self.cp, "ep": 1, "dp": self.dp, "pp": self.pp}
        return AxisLayout(tuple((n, sz[n]) for n in names))

    @property
    def expert(self) -> AxisLayout:
        names = tuple(n for n in _parse_order(self.order) if n != "cp")
        sz = {"tp": self.etp, "ep": self.ep, "dp": self.edp, "pp": self.pp}
        return AxisLayout(tuple((n, sz[n]) for n in names))

    # ---- collapsed MegaFBD: explicit rank arithmetic (blocks are not a regular grid) -----------
    def _cblock(self, p: int, r: int) -> int:
        return (p * self.dp + r) * (self.tp + 1)

    def _cfwd(self, p: int, r: int) -> int:
        return self._cblock(p, r)

    def _cbwd(self, p: int, r: int, t: int) -> int:
        return self._cblock(p, r) + 1 + t

    def _collapsed_coords(self, rank: int) -> Dict[str, int]:
        b, o = divmod(rank, self.tp + 1)
        p, r = divmod(b, self.dp)
        t = max(0, o - 1)
        return {"tp": t, "pp": p, "dp": r, "cp": 0, "ep": 0, "etp": t, "edp": r, "fb": 0 if o == 0 else 1, "hdp": r}

    def _collapsed_groups(self, kind: str) -> List[List[int]]:
        P, R, T = self.pp, self.dp, self.tp
        fw = lambda p, r: [self._cfwd(p, r)]                                    # noqa: E731
        bw = lambda p, r: [self._cbwd(p, r, t) for t in range(T)]               # noqa: E731
        if kind in ("tp", "etp", "tp_cp", "ep_tp"):
            return [g for p in range(P) for r in range(R) for g in (fw(p, r), bw(p, r))]
        if kind == "pp":
            return ([[self._cfwd(p, r) for p in range(P)] for r in range(R)]
                    + [[self._cbwd(p, r, t) for p in range(P)] for r in range(R) for t in range(T)])
        if kind in ("dp", "dp_cp", "edp", "half_dp", "intra_dp_cp"):
            return ([[self._cfwd(p, r) for r in range(R)] for p in range(P)]
                    + [[self._cbwd(p, r, t) for r in range(R)] for p in range(P) for t in range(T)])
        if kind in ("cp", "ep", "inter_dp_cp"):
            return [[k] for k in range(self.world_size)]
        if kind in ("mp", "tp_ep_pp"):
            return ([[self._cfwd(p, r) for p in range(P)] for r in range(R)]
                    + [[self._cbwd(p, r, t) for p in range(P) for t in range(T)] for r in range(R)])
        if kind in ("tp_dp_cp", "tp_dp"):
            return ([[self._cfwd(p, r) for r in range(R)] for p in range(P)]
                    + [[self._cbwd(p, r, t) for r in range(R) for t in range(T)] for p in range(P)])
        if kind == "embd":
            return [self.embedding_ranks(g) for g in self._collapsed_groups("pp")]
        if kind == "pos_embd":
            return [self.position_embedding_ranks(g) for g in self._collapsed_groups("pp")]
        if kind == "fb":          # [forward rank, backward tp rank 0, ..., tp - 1]
            return [fw(p, r) + bw(p, r) for p in range(P) for r in range(R)]
        raise KeyError(kind)

    # ---- coordinates ---------------------------------------------------
    def coords(self, rank: int) -> Dict[str, int]:
        if self.fbd_collapse:
            return self._collapsed_coords(rank)
        c = self.dense.coords(rank)
        e = self.expert.coords(rank)
        c["ep"] = e["ep"]
        c["etp"] = e["tp"]
        c["edp"] = e["dp"]
        if self.fbd:
            c["fb"] = c["dp"] % 2  # 0 = forward rank, 1 = backward rank
            c["hdp"] = c["dp"] // 2
        return c

    # ---- group families ------------------------------------------------
    def groups(self, kind: str) -> List[List[int]]:
        if self.fbd_collapse:
            return self._collapsed_groups(kind)
        d, e = self.dense, self.expert
        if kind == "tp":
            return d.groups("tp")
        if kind == "pp":
            return d.groups("pp")
        if kind == "dp":
            return d.groups("dp")
        if kind == "cp":
            return d.groups("cp")
        if kind == "dp_cp":
            return d.groups("dp-cp")
        if kind == "mp":  # model parallel = tp x pp
            return d.groups("tp-pp")
        if kind == "tp_cp":
            return d.groups("tp-cp")
        if kind == "tp_dp_cp":
            return d.groups("tp-dp-cp")
        if kind == "tp_dp":
            return d.groups("tp-dp")
        if kind == "ep":
            return e.groups("ep")
        if kind == "etp":
            return e.groups("tp")
        if kind == "ep_tp":
            return e.groups("tp-ep")
        if kind == "edp":
            return e.groups("dp")
        if kind == "tp_ep_pp":
            return e.groups("tp-ep-pp")
        if kind == "embd":
            return [self.embedding_ranks(g) for g in d.groups("pp")]
        if kind == "pos_embd":
            return [self.position_embedding_ranks(g) for g in d.groups("pp")]
        if kind == "fb":  # forward/backward twins (size-2 groups inside each dp group)
            out = []
            for g in d.groups("dp"):
                out += [[g[i], g[i + 1]] for i in range(0, len(g), 2)]
            return out
        if kind == "half_dp":  # every-other dp rank (same fb role)
            out = []
            for g in d.groups("dp"):
                out += [g[0::2], g[1::2]]
            return out
        if kind in ("hcp_inner", "hcp_outer"):
            if not self.hcp_sizes:
                return [[r] for r in range(self.world_size)]
            a = int(self.hcp_sizes[0])
            out = []
            for g in d.groups("cp"):
                if kind == "hcp_inner":
                    out += [g[i:i + a] for i in range(0, len(g), a)]
                else:
                    out += [g[i::a] for i in range(a)]
            return out
        if kind == "intra_dp_cp":
            n = self.num_distributed_optimizer_instances
            out = []
            for g in d.groups("dp-cp"):
                chunk = len(g) // n
                out += [g[i * chunk:(i + 1) * chunk] for i in range(n)]
            return out
        if kind == "inter_dp_cp":
            n = self.num_distributed_optimizer_instances
            out = []
            for g in d.groups("dp-cp"):
                chunk = len(g) // n
                out += [g[i::chunk] for i in range(chunk)]
            return out
        raise KeyError(kind)

    def group_of(self, kind: str, rank: int) -> List[int]:
        for g in self.groups(kind):
            if rank in g:
                return g
        raise ValueError(f"rank {rank} is in no {kind} group")

    def embedding_ranks(self, pp_group: Sequence[int]) -> List[int]:
        """First and last pipeline stage share the tied embedding gradient; with an encoder-decoder split the
        first decoder stage holds a third copy (reference ``parallel_state.default_embedding_ranks``)."""
        if len(pp_group) == 1:
            return [pp_group[0]]
        out = [pp_group[0], pp_group[-1]]
        if self.split_rank is not None and pp_group[self.split_rank] not in out:
            out.insert(1, pp_group[self.split_rank])
        return out

    def position_embedding_ranks(self, pp_group: Sequence[int]) -> List[int]:
        out = [pp_group[0]]
        if self.split_rank is not None and len(pp_group) > 1 and pp_group[self.split_rank] not in out:
            out.append(pp_group[self.split_rank])
        return out

    # ---- convenience ---------------------------------------------------
    @property
    def effective_dp(self) -> int:
        if self.fbd_collapse:
            return self.dp
        return self.dp // 2 if self.fbd else self.dp

    def is_forward_rank(self, rank: int) -> bool:
        if self.fbd_collapse:
            return rank % (self.tp + 1) == 0
        return self.fbd and self.coords(rank)["dp"] % 2 == 0

    def is_backward_rank(self, rank: int) -> bool:
        if self.fbd_collapse:
            return rank % (self.tp + 1) != 0
        return (not self.fbd) or self.coords(rank)["dp"] % 2 == 1

    def twin_of(self, rank: int) -> int:
        """The partner across the forward/backward split; under the collapsed layout a backward rank's twin is its
        block's forward rank and a forward rank's "twin" is backward TP rank 0 (use ``group_of('fb')`` for all)."""
        if not self.fbd:
            raise ValueError("no twin without fbd")
        g = self.group_of("fb", rank)
        return g[1] if g[0] == rank else g[0]

    def describe(self) -> str:
        return (
            f"Topology(world={self.world_size}, tp={self.tp}, pp={self.pp}, cp={self.cp}, "
            f"dp={self.dp}, ep={self.ep}, etp={self.etp}, edp={self.edp}, vpp={self.vpp}, "
            f"fbd={self.fbd}, order={self.order})"
        )


def layer_offset(num_layers: int, pp: int, pp_rank: int, vpp: int | None = None,
                 vpp_rank: int | None = None, first_stage_layers: int | None = None,
                 last_stage_layers: int | None = None) -> Tuple[int, int]:
    """(offset, count) of the transformer layers owned by a (pp_rank, vpp_rank) chunk.

    Interleaved assignment: with ``vpp`` chunks per stage, chunk ``v`` of stage ``p``
    owns layers ``[(v*pp + p) * L, ...)`` with ``L = num_layers / (pp*vpp)``.
    Parity: reference ``transformer_layer.py get_transformer_layer_offset``.
    """
    if first_stage_layers is not None or last_stage_layers is not None:
        if vpp not in (None, 1):
            raise ValueError("uneven pipeline split is not combined with vpp")
        first = first_stage_layers
        last = last_stage_layers
        middle_stages = pp - (first is not None) - (last is not None)
        middle_layers = num_layers - (first or 0) - (last or 0)
        per = middle_layers // middle_stages if middle_stages else 0
        counts = []
        for p in range(pp):
            if p == 0 and first is not None:
                counts.append(first)
            elif p == pp - 1 and last is not None:
                counts.append(last)
            else:
                counts.append(per)
        if sum(counts) != num_layers:
            raise ValueError("uneven split does not add up to num_layers")
        return sum(counts[:pp_rank]), counts[pp_rank]
    v = 1 if vpp is None else vpp
    if num_layers % (pp * v) != 0:
        raise ValueError(f"num_layers {num_layers} not divisible by pp*vpp {pp * v}")
    per = num_layers // (pp * v)
    vr = 0 if vpp_rank is None else vpp_rank
    return (vr * pp + pp_rank) * per, per
