"""THD packed-sequence descriptor (reference ``core/packed_seq_params.py``)."""
from dataclasses import dataclass
from typing import Optional

import torch


@dataclass
class PackedSeqParams:
    qkv_format: str = None
    cu_seqlens_q: Optional[torch.Tensor] = None
    cu_seqlens_kv: Optional[torch.Tensor] = None
    cu_seqlens_q_padded: Optional[torch.Tensor] = None
    cu_seqlens_kv_padded: Optional[torch.Tensor] = None
    max_seqlen_q: Optional[int] = None
    max_seqlen_kv: Optional[int] = None
