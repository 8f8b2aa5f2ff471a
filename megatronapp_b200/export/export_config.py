"""Parallel layout of the exported checkpoint (reference ``core/export/export_config.py``)."""
from dataclasses import dataclass
from typing import Optional


@dataclass
class ExportConfig:
    inference_tp_size: int = 1
    inference_pp_size: int = 1
    use_parallel_embedding: bool = False
    use_embedding_sharing: Optional[bool] = None          # deprecated in the reference: taken from the model instead

    def __post_init__(self):
        if self.inference_tp_size < 1 or self.inference_pp_size < 1:
            raise ValueError("inference_tp_size and inference_pp_size must be >= 1")
