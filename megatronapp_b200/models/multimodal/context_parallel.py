"""Sequence / context-parallel padding of the combined text + image sequence (reference
``core/models/multimodal/context_parallel.py``)."""
from __future__ import annotations

import torch

from ...utils.packed_seq_params import PackedSeqParams


def get_padding(seq_len: int, cp_size: int, tp_size: int, has_sp: bool, decoder_tp_comm_overlap: bool = False,
                decoder_seq_len: int = None) -> int:
    """Tokens to append so the combined sequence divides evenly: by ``tp`` for sequence parallelism, by ``2 * cp`` for the
    zig-zag context-parallel split, by ``2 * cp * tp`` for both; with the fused TP collectives on the language model the
    buffers are sized for ``decoder_seq_len``, so the sequence is padded to exactly that."""
    if has_sp and decoder_tp_comm_overlap:
        if decoder_seq_len is None:
            raise ValueError("decoder_seq_len is required when the language model uses fused TP communication")
        return decoder_seq_len - seq_len
    factor = 1
    if has_sp and cp_size > 1:
        factor = tp_size * cp_size * 2
    elif cp_size > 1:
        factor = cp_size * 2
    elif has_sp:
        factor = tp_size
    return (seq_len + factor - 1) // factor * factor - seq_len


def get_packed_seq_params(tokens: torch.Tensor, img_seq_len: int, padding_needed: int, cp_size: int,
                          use_packed_sequence: bool = False) -> PackedSeqParams:
    """Sequence boundaries of the combined batch: ``cu_seqlens`` count the valid tokens of every sample, the ``_padded``
    variant (only with context parallelism + padding, where the THD format is required) the padded lengths."""
    b = tokens.shape[0]
    valid = tokens.shape[1] + img_seq_len - padding_needed
    padded = tokens.shape[1] + img_seq_len
    cu = torch.arange(0, (b + 1) * valid, step=valid, dtype=torch.int32, device=tokens.device)
    cu_padded, fmt = None, "sbhd"
    if cp_size > 1 and (padding_needed > 0 or use_packed_sequence):
        cu_padded = torch.arange(0, (b + 1) * padded, step=padded, dtype=torch.int32, device=tokens.device)
        fmt = "thd"
    return PackedSeqParams(qkv_format=fmt, cu_seqlens_q=cu, cu_seqlens_kv=cu, cu_seqlens_q_padded=cu_padded,
                           cu_seqlens_kv_padded=cu_padded, max_seqlen_q=padded, max_seqlen_kv=padded)
