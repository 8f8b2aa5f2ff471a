"""THD packed-sequence descriptor (reference ``core/packed_seq_params.py``)."""
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

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
    # host copies of the boundaries: read ONCE per packed batch (the first attention layer), reused by every other layer --
    # the per-sequence kernel launches need Python ints, and a .tolist() per layer is a device sync per layer.  Pass CPU
    # tensors / lists from the data pipeline and no sync happens at all.
    _host: Optional[Tuple[List[int], List[int]]] = field(default=None, repr=False, compare=False)

    def boundaries(self) -> Tuple[List[int], List[int]]:
        if self._host is None:
            def to_list(t):
                return [int(x) for x in (t.tolist() if torch.is_tensor(t) else t)]

            q = to_list(self.cu_seqlens_q)
            self._host = (q, q if self.cu_seqlens_kv is None else to_list(self.cu_seqlens_kv))
        return self._host
