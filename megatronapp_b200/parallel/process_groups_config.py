"""Explicit process-group bundle (reference ``core/process_groups_config.py``): lets a module be built
against caller-supplied groups instead of the global parallel state."""
from dataclasses import dataclass, fields
from typing import Any, Optional


@dataclass
class ModelCommProcessGroups:
    tp: Optional[Any] = None
    pp: Optional[Any] = None
    mp: Optional[Any] = None
    embd: Optional[Any] = None
    pos_embd: Optional[Any] = None
    cp: Optional[Any] = None
    tp_cp: Optional[Any] = None
    hcp: Optional[Any] = None
    ep: Optional[Any] = None
    expt_tp: Optional[Any] = None
    tp_ep: Optional[Any] = None
    tp_ep_pp: Optional[Any] = None
    dp: Optional[Any] = None
    dp_cp: Optional[Any] = None
    expt_dp: Optional[Any] = None

    @classmethod
    def use_mpu_process_groups(cls, required_pgs=None):
        from . import state as ps

        getters = {"tp": ps.get_tensor_model_parallel_group, "pp": ps.get_pipeline_model_parallel_group,
                   "mp": ps.get_model_parallel_group, "embd": ps.get_embedding_group,
                   "cp": ps.get_context_parallel_group, "tp_cp": ps.get_tensor_and_context_parallel_group,
                   "ep": ps.get_expert_model_parallel_group, "dp": ps.get_data_parallel_group,
                   "dp_cp": lambda: ps.get_data_parallel_group(with_context_parallel=True)}
        names = required_pgs or [f.name for f in fields(cls) if f.name in getters]
        return cls(**{n: getters[n]() for n in names if n in getters})


@dataclass
class GradCommProcessGroups:
    dp: Optional[Any] = None
    dp_cp: Optional[Any] = None
    expt_dp: Optional[Any] = None
    intra_dp_cp: Optional[Any] = None
    inter_dist_opt: Optional[Any] = None
