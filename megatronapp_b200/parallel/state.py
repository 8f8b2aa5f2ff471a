"""Process-group registry with the Megatron ``parallel_state`` API surface.

One :class:`ParallelState` instance owns every ``torch.distributed`` group for a
:class:`~megatronapp_b200.parallel.topology.Topology`.  Module-level functions
(``initialize_model_parallel``, ``get_tensor_model_parallel_group`` ...) operate
on the current instance so user code written against
``megatron.core.parallel_state`` keeps working.

Parity: reference ``megatron/core/parallel_state.py:1272-2056`` (group creation),
``:2058-3180`` (accessors), ``:3180`` (destroy); MegaFBD groups ``:444-512``.
Differences by design: groups come from a pure ``Topology``; there are no
per-thread "virtual" group lists and no coordinator rank (see ``fbd/``).
"""
from __future__ import annotations

import datetime
import os
from typing import Dict, List, Optional

import torch
import torch.distributed as dist

from .topology import Topology

_STATE: Optional["ParallelState"] = None


class _SingleGroup:
    """Stand-in group for world_size == 1 without an initialized backend."""

    def size(self):
        return 1

    def rank(self):
        return 0


class ParallelState:
    def __init__(self, topo: Topology, rank: int, backend: Optional[str] = None,
                 timeout_minutes: int = 30, create_gloo_process_groups: bool = False,
                 nccl_options: Optional[Dict[str, dict]] = None):
        self.topo = topo
        self.rank = rank
        self.coords = topo.coords(rank)
        self.groups: Dict[str, object] = {}
        self.group_ranks: Dict[str, List[int]] = {}
        self.gloo_groups: Dict[str, object] = {}
        self.virtual_pp_rank: Optional[int] = None
        self.vpp = topo.vpp
        self._memory_buffers: Dict[tuple, torch.Tensor] = {}
        self.distributed = dist.is_available() and dist.is_initialized()
        if topo.world_size > 1 and not self.distributed:
            raise RuntimeError("torch.distributed must be initialized for world_size > 1")
        timeout = datetime.timedelta(minutes=timeout_minutes)
        kinds = ["tp", "pp", "dp", "cp", "dp_cp", "mp", "tp_cp", "tp_dp_cp", "tp_dp", "ep",
                 "etp", "ep_tp", "edp", "tp_ep_pp", "embd", "pos_embd"]
        if topo.num_distributed_optimizer_instances > 1:
            kinds += ["intra_dp_cp", "inter_dp_cp"]
        if topo.fbd:
            kinds += ["fb", "half_dp"]
        if topo.hcp_sizes:
            kinds += ["hcp_inner", "hcp_outer"]
        for kind in kinds:
            all_groups = topo.groups(kind)
            mine = None
            for g in all_groups:
                pg = None
                if self.distributed and len(g) > 1:
                    opts = None
                    if nccl_options and kind in nccl_options and backend in (None, "nccl"):
                        opts = _make_nccl_options(nccl_options[kind])
                    pg = dist.new_group(ranks=g, timeout=timeout, pg_options=opts)
                if rank in g:
                    mine = g
                    self.groups[kind] = pg if pg is not None else _SingleGroup()
                    if (create_gloo_process_groups and self.distributed and len(g) > 1
                            and kind in ("dp", "dp_cp", "edp")):
                        pass  # created below (needs collective ordering across all ranks)
            if mine is None:
                # e.g. embedding group for a middle pipeline stage
                self.groups[kind] = None
                self.group_ranks[kind] = []
            else:
                self.group_ranks[kind] = mine
        if create_gloo_process_groups and self.distributed:
            for kind in ("dp", "dp_cp", "edp"):
                for g in topo.groups(kind):
                    if len(g) > 1:
                        pg = dist.new_group(ranks=g, timeout=timeout, backend="gloo")
                        if rank in g:
                            self.gloo_groups[kind] = pg

    # -- generic accessors ------------------------------------------------
    def group(self, kind: str):
        return self.groups[kind]

    def ranks(self, kind: str) -> List[int]:
        return self.group_ranks[kind]

    def world(self, kind: str) -> int:
        return max(1, len(self.group_ranks[kind]))

    def rank_in(self, kind: str) -> int:
        r = self.group_ranks[kind]
        return r.index(self.rank) if self.rank in r else 0


def _make_nccl_options(cfg: dict):
    opts = dist.ProcessGroupNCCL.Options()
    if "cga_cluster_size" in cfg:
        opts.config.cga_cluster_size = cfg["cga_cluster_size"]
    if "max_ctas" in cfg:
        opts.config.max_ctas = cfg["max_ctas"]
    if "min_ctas" in cfg:
        opts.config.min_ctas = cfg["min_ctas"]
    if cfg.get("is_high_priority_stream"):
        opts.is_high_priority_stream = True
    return opts


# ---------------------------------------------------------------------------
# Megatron-compatible module-level API
# ---------------------------------------------------------------------------

def initialize_model_parallel(
    tensor_model_parallel_size: int = 1,
    pipeline_model_parallel_size: int = 1,
    virtual_pipeline_model_parallel_size: Optional[int] = None,
    pipeline_model_parallel_split_rank: Optional[int] = None,
    pipeline_model_parallel_comm_backend: Optional[str] = None,
    use_sharp: bool = False,
    context_parallel_size: int = 1,
    hierarchical_context_parallel_sizes: Optional[List[int]] = None,
    expert_model_parallel_size: int = 1,
    num_distributed_optimizer_instances: int = 1,
    expert_tensor_parallel_size: Optional[int] = None,
    nccl_communicator_config_path: Optional[str] = None,
    distributed_timeout_minutes: int = 30,
    order: str = "tp-cp-ep-dp-pp",
    encoder_tensor_model_parallel_size: int = 0,
    encoder_pipeline_model_parallel_size: Optional[int] = 0,
    get_embedding_ranks=None,
    get_position_embedding_ranks=None,
    create_gloo_process_groups: bool = False,
    forward_backward_disaggregating: bool = False,
    ignore_forward_tensor_parallel: bool = False,
    world_size: Optional[int] = None,
    rank: Optional[int] = None,
) -> ParallelState:
    global _STATE
    if _STATE is not None:
        raise RuntimeError("model parallel is already initialized")
    if dist.is_available() and dist.is_initialized():
        ws, rk = dist.get_world_size(), dist.get_rank()
    else:
        ws, rk = (world_size or 1), (rank or 0)
    nccl_options = None
    if nccl_communicator_config_path:
        import yaml

        with open(nccl_communicator_config_path) as f:
            nccl_options = yaml.safe_load(f)
    topo = Topology(
        world_size=ws,
        tp=tensor_model_parallel_size,
        pp=pipeline_model_parallel_size,
        cp=context_parallel_size,
        ep=expert_model_parallel_size,
        etp=expert_tensor_parallel_size,
        vpp=virtual_pipeline_model_parallel_size,
        order=order,
        fbd=forward_backward_disaggregating,
        fbd_collapse=bool(forward_backward_disaggregating and ignore_forward_tensor_parallel),
        num_distributed_optimizer_instances=num_distributed_optimizer_instances,
        split_rank=(encoder_pipeline_model_parallel_size or pipeline_model_parallel_split_rank or None)
        if pipeline_model_parallel_size > 1 else None,
        hcp_sizes=tuple(int(x) for x in hierarchical_context_parallel_sizes) if hierarchical_context_parallel_sizes else None,
    )
    backend = dist.get_backend() if (dist.is_available() and dist.is_initialized()) else None
    _STATE = ParallelState(topo, rk, backend=backend,
                           timeout_minutes=distributed_timeout_minutes,
                           create_gloo_process_groups=create_gloo_process_groups,
                           nccl_options=nccl_options)
    if virtual_pipeline_model_parallel_size is not None:
        _STATE.virtual_pp_rank = 0
    return _STATE


def model_parallel_is_initialized() -> bool:
    return _STATE is not None


def is_initialized() -> bool:
    return _STATE is not None


def is_unitialized() -> bool:  # (sic) name kept for API compatibility
    return _STATE is None


def get_state() -> ParallelState:
    if _STATE is None:
        raise RuntimeError("model parallel is not initialized")
    return _STATE


def get_topology() -> Topology:
    return get_state().topo


def destroy_model_parallel() -> None:
    global _STATE
    if _STATE is not None and _STATE.distributed:
        for pg in list(_STATE.groups.values()) + list(_STATE.gloo_groups.values()):
            if pg is not None and not isinstance(pg, _SingleGroup):
                try:
                    dist.destroy_process_group(pg)
                except Exception:
                    pass
    _STATE = None


def _acc(kind):
    def group(check_initialized: bool = True, **_kw):
        if _STATE is None:
            if check_initialized:
                raise RuntimeError(f"{kind} group is not initialized")
            return None
        return _STATE.group(kind)

    def world(**_kw):
        return 1 if _STATE is None else _STATE.world(kind)

    def rank(**_kw):
        return 0 if _STATE is None else _STATE.rank_in(kind)

    return group, world, rank


get_tensor_model_parallel_group, get_tensor_model_parallel_world_size, get_tensor_model_parallel_rank = _acc("tp")
get_pipeline_model_parallel_group, get_pipeline_model_parallel_world_size, get_pipeline_model_parallel_rank = _acc("pp")
get_context_parallel_group, get_context_parallel_world_size, get_context_parallel_rank = _acc("cp")
get_model_parallel_group, get_model_parallel_world_size, get_model_parallel_rank = _acc("mp")
get_expert_model_parallel_group, get_expert_model_parallel_world_size, get_expert_model_parallel_rank = _acc("ep")
get_expert_tensor_parallel_group, get_expert_tensor_parallel_world_size, get_expert_tensor_parallel_rank = _acc("etp")
get_expert_tensor_and_model_parallel_group, get_expert_tensor_and_model_parallel_world_size, get_expert_tensor_and_model_parallel_rank = _acc("ep_tp")
get_expert_data_parallel_group, get_expert_data_parallel_world_size, get_expert_data_parallel_rank = _acc("edp")
get_embedding_group, _embd_world, _embd_rank = _acc("embd")
get_position_embedding_group, _pe_world, _pe_rank = _acc("pos_embd")
get_forward_backward_parallel_group, get_forward_backward_parallel_world_size, get_forward_backward_parallel_rank = _acc("fb")
get_tensor_and_context_parallel_group, get_tensor_and_context_parallel_world_size, get_tensor_and_context_parallel_rank = _acc("tp_cp")


def get_hierarchical_context_parallel_groups(check_initialized: bool = True):
    """``[inner (all-to-all) group, outer (ring) group]`` of this rank (reference ``parallel_state.py`` ``_HIERARCHICAL_CONTEXT_PARALLEL_GROUPS``);
    ``None`` when ``--hierarchical-context-parallel-sizes`` was not given."""
    st = get_state()
    if not st.topo.hcp_sizes:
        return None
    return [st.group("hcp_inner"), st.group("hcp_outer")]


def _dp_kind(with_context_parallel: bool = False) -> str:
    st = get_state()
    if st.topo.fbd:
        return "half_dp"
    return "dp_cp" if with_context_parallel else "dp"


def get_data_parallel_group(with_context_parallel: bool = False, partial_data_parallel: bool = False):
    st = get_state()
    if partial_data_parallel and st.topo.num_distributed_optimizer_instances > 1:
        return st.group("intra_dp_cp")
    return st.group(_dp_kind(with_context_parallel))


def get_data_parallel_group_gloo(with_context_parallel: bool = False, **_kw):
    st = get_state()
    return st.gloo_groups.get("dp_cp" if with_context_parallel else "dp")


def get_data_parallel_world_size(with_context_parallel: bool = False, **_kw) -> int:
    if _STATE is None:
        return 1
    return _STATE.world(_dp_kind(with_context_parallel))


def get_data_parallel_rank(with_context_parallel: bool = False, **_kw) -> int:
    if _STATE is None:
        return 0
    return _STATE.rank_in(_dp_kind(with_context_parallel))


def get_full_data_parallel_group():
    """All DP replicas, including both halves of every MegaFBD pair."""
    return get_state().group("dp")


def get_half_data_parallel_group():
    return get_state().group("half_dp")


def get_tensor_and_data_parallel_group(with_context_parallel: bool = False):
    return get_state().group("tp_dp_cp" if with_context_parallel else "tp_dp")


def get_tensor_and_data_parallel_ranks(with_context_parallel: bool = False) -> List[int]:
    return get_state().ranks("tp_dp_cp" if with_context_parallel else "tp_dp")


def get_group_ranks(kind: str) -> List[int]:
    return get_state().ranks(kind)


# ---- pipeline helpers -------------------------------------------------------

def get_virtual_pipeline_model_parallel_world_size() -> Optional[int]:
    return None if _STATE is None else _STATE.vpp


def get_virtual_pipeline_model_parallel_rank() -> Optional[int]:
    return None if _STATE is None else _STATE.virtual_pp_rank


def set_virtual_pipeline_model_parallel_rank(rank: Optional[int]) -> None:
    get_state().virtual_pp_rank = rank


def set_virtual_pipeline_model_parallel_world_size(ws: Optional[int]) -> None:
    get_state().vpp = ws


def is_pipeline_first_stage(ignore_virtual: bool = False) -> bool:
    if _STATE is None:
        return True
    if not ignore_virtual and _STATE.vpp is not None and _STATE.virtual_pp_rank not in (None, 0):
        return False
    return get_pipeline_model_parallel_rank() == 0


def is_pipeline_last_stage(ignore_virtual: bool = False) -> bool:
    if _STATE is None:
        return True
    if (not ignore_virtual and _STATE.vpp is not None and _STATE.virtual_pp_rank is not None
            and _STATE.virtual_pp_rank != _STATE.vpp - 1):
        return False
    return get_pipeline_model_parallel_rank() == get_pipeline_model_parallel_world_size() - 1


def get_pipeline_model_parallel_first_rank() -> int:
    return get_state().ranks("pp")[0]


def get_pipeline_model_parallel_last_rank() -> int:
    return get_state().ranks("pp")[-1]


def get_pipeline_model_parallel_next_rank() -> int:
    r = get_state().ranks("pp")
    return r[(get_pipeline_model_parallel_rank() + 1) % len(r)]


def get_pipeline_model_parallel_prev_rank() -> int:
    r = get_state().ranks("pp")
    return r[(get_pipeline_model_parallel_rank() - 1) % len(r)]


def get_tensor_model_parallel_src_rank() -> int:
    return get_state().ranks("tp")[0]


def get_data_parallel_src_rank(with_context_parallel: bool = False) -> int:
    return get_state().ranks(_dp_kind(with_context_parallel))[0]


def is_rank_in_embedding_group(ignore_virtual: bool = False) -> bool:
    if _STATE is None:
        return True
    return _STATE.rank in _STATE.ranks("embd")


# ---- encoder / decoder pipeline split (T5-style models; reference parallel_state.is_inside_encoder ...) --------
def get_pipeline_model_parallel_split_rank() -> Optional[int]:
    return None if _STATE is None else _STATE.topo.split_rank


def is_inside_encoder(rank: Optional[int] = None) -> bool:
    split = get_pipeline_model_parallel_split_rank()
    if split is None:
        return True
    return (get_pipeline_model_parallel_rank() if rank is None else rank) < split


def is_inside_decoder(rank: Optional[int] = None) -> bool:
    split = get_pipeline_model_parallel_split_rank()
    if split is None:
        return True
    return (get_pipeline_model_parallel_rank() if rank is None else rank) >= split


def is_pipeline_stage_before_split(rank: Optional[int] = None) -> bool:
    return get_pipeline_model_parallel_world_size() == 1 or is_inside_encoder(rank) or \
        get_pipeline_model_parallel_split_rank() is None


def is_pipeline_stage_after_split(rank: Optional[int] = None) -> bool:
    return get_pipeline_model_parallel_world_size() == 1 or is_inside_decoder(rank) or \
        get_pipeline_model_parallel_split_rank() is None


def is_pipeline_stage_at_split() -> bool:
    r = get_pipeline_model_parallel_rank()
    return is_pipeline_stage_before_split(r) and is_pipeline_stage_after_split(r + 1)


def is_rank_in_position_embedding_group() -> bool:
    if _STATE is None:
        return True
    return _STATE.rank in _STATE.ranks("pos_embd")


# ---- MegaFBD ---------------------------------------------------------------

def is_forward_stage() -> bool:
    """True on a MegaFBD forward-only rank."""
    return _STATE is not None and _STATE.topo.fbd and _STATE.topo.is_forward_rank(_STATE.rank)


def is_backward_stage() -> bool:
    return _STATE is None or (not _STATE.topo.fbd) or _STATE.topo.is_backward_rank(_STATE.rank)


def fbd_forward_tp_collapsed() -> bool:
    """``--ignore-forward-tensor-parallel``: one un-sharded forward rank serves the tp backward ranks of its stage."""
    return _STATE is not None and _STATE.topo.fbd_collapse


def get_fbd_block_ranks() -> List[int]:
    """[forward rank, backward tp rank 0, ..., tp - 1] of this rank's (stage, replica) block."""
    return get_state().ranks("fb")


def get_twin_rank() -> int:
    st = get_state()
    return st.topo.twin_of(st.rank)


# ---- global scratch ----------------------------------------------------------

class GlobalMemoryBuffer:
    """Named, size-keyed scratch tensors reused across layers.

    Parity: reference ``core/utils.py GlobalMemoryBuffer`` / ``parallel_state.py:3142``.
    """

    def __init__(self):
        self.buffer: Dict[tuple, torch.Tensor] = {}

    def get_tensor(self, tensor_shape, dtype, name, device=None):
        numel = 1
        for s in tensor_shape:
            numel *= s
        key = (name, dtype, str(device))
        buf = self.buffer.get(key)
        if buf is None or buf.numel() < numel:
            dev = device if device is not None else (
                torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else "cpu")
            buf = torch.empty(numel, dtype=dtype, device=dev, requires_grad=False)
            self.buffer[key] = buf
        return buf[:numel].view(*tensor_shape)


_GLOBAL_MEMORY_BUFFER = GlobalMemoryBuffer()


def get_global_memory_buffer() -> GlobalMemoryBuffer:
    return _GLOBAL_MEMORY_BUFFER


def destroy_global_memory_buffer() -> None:
    global _GLOBAL_MEMORY_BUFFER
    _GLOBAL_MEMORY_BUFFER = GlobalMemoryBuffer()


def init_distributed(backend: Optional[str] = None, timeout_minutes: int = 30) -> None:
    """Initialize ``torch.distributed`` from torchrun-style env vars (one process per GPU)."""
    if dist.is_initialized():
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if torch.cuda.is_available():
        torch.cuda.set_device(local_rank % torch.cuda.device_count())
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29500")
    kwargs = {}
    if backend == "nccl":
        kwargs["device_id"] = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group(backend=backend, world_size=world, rank=rank,
                            timeout=datetime.timedelta(minutes=timeout_minutes), **kwargs)


# -- group-construction helpers of the reference's ``parallel_state.py`` -----------------------------------------------------------
def get_nccl_options(pg_name: str, nccl_comm_cfgs: dict):
    """``ProcessGroupNCCL.Options`` for ``pg_name`` from a communicator config (``--nccl-communicator-config-path`` yaml:
    ``{group: {cga_cluster_size, max_ctas, min_ctas, net_name}}``); ``None`` when the group has no entry."""
    if not nccl_comm_cfgs or pg_name not in nccl_comm_cfgs:
        return None
    cfg = nccl_comm_cfgs[pg_name]
    opts = dist.ProcessGroupNCCL.Options()
    opts.config.cga_cluster_size = cfg.get("cga_cluster_size", 4)
    opts.config.max_ctas = cfg.get("max_ctas", 32)
    opts.config.min_ctas = cfg.get("min_ctas", 1)
    if "net_name" in cfg:
        opts.config.net_name = cfg["net_name"]
    return opts


def create_group(ranks=None, timeout=None, backend=None, pg_options=None, use_local_synchronization: bool = False, group_desc=None):
    """``torch.distributed.new_group`` with the keyword set the reference uses (a description where torch supports one)."""
    kwargs = {"ranks": ranks, "timeout": timeout, "backend": backend, "pg_options": pg_options,
              "use_local_synchronization": use_local_synchronization}
    try:
        return dist.new_group(**kwargs, group_desc=group_desc)
    except TypeError:                      # older torch: no ``group_desc``
        return dist.new_group(**kwargs)


def generate_masked_orthogonal_rank_groups(world_size: int, parallel_size: List[int], mask: List[bool]) -> List[List[int]]:
    """All rank groups that vary along the masked dimensions of a ``parallel_size`` grid (first dimension fastest) while the
    unmasked coordinates stay fixed -- the rule ``RankGenerator`` derives every tensor / pipeline / data / context / expert
    parallel group from (reference ``parallel_state.py``)."""
    def prefix_product(a, init=1):
        r = [init]
        for v in a:
            init = init * v
            r.append(init)
        return r

    def decompose(index, shape):
        stride = prefix_product(shape)
        return [(index // d) % s for s, d in zip(shape, stride)]

    masked_shape = [s for s, m in zip(parallel_size, mask) if m]
    unmasked_shape = [s for s, m in zip(parallel_size, mask) if not m]
    global_stride = prefix_product(parallel_size)
    masked_stride = [d for d, m in zip(global_stride, mask) if m]
    unmasked_stride = [d for d, m in zip(global_stride, mask) if not m]
    group_size = prefix_product(masked_shape)[-1]
    num_of_group = world_size // group_size
    ranks = []
    for group_index in range(num_of_group):
        fixed = decompose(group_index, unmasked_shape)
        base = sum(x * y for x, y in zip(fixed, unmasked_stride))
        ranks.append([base + sum(x * y for x, y in zip(decompose(r, masked_shape), masked_stride)) for r in range(group_size)])
    return ranks
