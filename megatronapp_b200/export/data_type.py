"""Export precision (reference ``core/export/data_type.py``)."""
from enum import Enum

import torch


class DataType(Enum):
    bfloat16 = "bfloat16"
    float16 = "float16"
    float32 = "float32"

    @property
    def torch_dtype(self) -> torch.dtype:
        return {"bfloat16": torch.bfloat16, "float16": torch.float16, "float32": torch.float32}[self.value]

    @classmethod
    def of(cls, x) -> "DataType":
        if isinstance(x, cls):
            return x
        if isinstance(x, torch.dtype):
            return {torch.bfloat16: cls.bfloat16, torch.float16: cls.float16, torch.float32: cls.float32}[x]
        return cls(str(x))
