"""Token / position embeddings and rotary position embeddings.

Parity: reference ``models/common/embeddings/language_model_embedding.py`` (word + learned
absolute position (+ tokentype) embeddings, dropout, scatter to the sequence-parallel region),
``rotary_pos_embedding.py`` (``RotaryEmbedding`` incl. ``rotary_percent``, linear
``seq_len_interpolation_factor`` and llama-3 style ``rope_scaling``) and
``yarn_rotary_pos_embedding.py``.  The rotary table is a fp32 ``[s, rot_dim]`` (cos | sin)
tensor consumed by the fused RoPE kernel.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn

from ...parallel import state as ps
from ...parallel.mappings import scatter_to_sequence_parallel_region
from ...parallel.random import get_cuda_rng_tracker
from ...parallel.tp_layers import VocabParallelEmbedding
from ...transformer.module import MegatronModule


class LanguageModelEmbedding(MegatronModule):
    def __init__(self, config, vocab_size: int, max_sequence_length: int,
                 position_embedding_type: str = "learned_absolute", num_tokentypes: int = 0,
                 scatter_to_sequence_parallel: bool = True):
        super().__init__(config)
        self.vocab_size, self.max_sequence_length = vocab_size, max_sequence_length
        self.add_position_embedding = position_embedding_type == "learned_absolute"
        self.num_tokentypes = num_tokentypes
        self.scatter_to_sequence_parallel = scatter_to_sequence_parallel
        self.reduce_scatter_embeddings = ((not self.add_position_embedding) and num_tokentypes <= 0
                                          and config.sequence_parallel and scatter_to_sequence_parallel)
        self.word_embeddings = VocabParallelEmbedding(
            vocab_size, config.hidden_size, init_method=config.init_method,
            reduce_scatter_embeddings=self.reduce_scatter_embeddings, config=config)
        dev = "cpu" if (config.use_cpu_initialization or not torch.cuda.is_available()) else torch.cuda.current_device()
        if self.add_position_embedding:
            self.position_embeddings = nn.Embedding(max_sequence_length, config.hidden_size,
                                                    dtype=config.params_dtype, device=dev)
            if config.perform_initialization:
                config.init_method(self.position_embeddings.weight)
        if num_tokentypes > 0:
            self.tokentype_embeddings = nn.Embedding(num_tokentypes, config.hidden_size,
                                                     dtype=config.params_dtype, device=dev)
            if config.perform_initialization:
                config.init_method(self.tokentype_embeddings.weight)
        else:
            self.tokentype_embeddings = None
        self.embedding_dropout = nn.Dropout(config.hidden_dropout)

    def zero_parameters(self):
        for p in self.parameters():
            p.data.fill_(0)
            p.shared = True

    def forward(self, input_ids, position_ids, tokentype_ids=None):
        emb = self.word_embeddings(input_ids)
        if self.add_position_embedding:
            emb = emb + self.position_embeddings(position_ids)
        if not self.reduce_scatter_embeddings:
            emb = emb.transpose(0, 1).contiguous()  # [b, s, h] -> [s, b, h]
        if tokentype_ids is not None and self.tokentype_embeddings is not None:
            emb = emb + self.tokentype_embeddings(tokentype_ids).permute(1, 0, 2)
        if self.config.fp32_residual_connection:
            emb = emb.float()
        if self.config.sequence_parallel:
            if not self.reduce_scatter_embeddings and self.scatter_to_sequence_parallel:
                emb = scatter_to_sequence_parallel_region(emb)
            if self.config.clone_scatter_output_in_embedding and self.scatter_to_sequence_parallel:
                emb = emb.clone()
            tr = get_cuda_rng_tracker()
            if "model-parallel-rng" in tr.states_ and self.config.hidden_dropout > 0:
                with tr.fork():
                    emb = self.embedding_dropout(emb)
            else:
                emb = self.embedding_dropout(emb)
        else:
            emb = self.embedding_dropout(emb)
        return emb


def _apply_llama3_scaling(inv_freq, factor=8.0, low_freq_factor=1.0, high_freq_factor=4.0, original_max=8192):
    low_wl = original_max / low_freq_factor
    high_wl = original_max / high_freq_factor
    wl = 2 * math.pi / inv_freq
    scaled = torch.where(wl > low_wl, inv_freq / factor, inv_freq)
    smooth = (original_max / wl - low_freq_factor) / (high_freq_factor - low_freq_factor)
    smoothed = (1 - smooth) * inv_freq / factor + smooth * inv_freq
    mid = (wl >= high_wl) & (wl <= low_wl)
    return torch.where(mid, smoothed, scaled)


class RotaryEmbedding(nn.Module):
    def __init__(self, kv_channels: int, rotary_percent: float = 1.0, rotary_interleaved: bool = False,
                 seq_len_interpolation_factor: Optional[float] = None, rotary_base: float = 10000.0,
                 rope_scaling: bool = False, rope_scaling_factor: float = 8.0, use_cpu_initialization: bool = False):
        super().__init__()
        dim = kv_channels
        if rotary_percent < 1.0:
            dim = int(dim * rotary_percent)
        self.dim = dim
        self.rotary_interleaved = rotary_interleaved
        self.seq_len_interpolation_factor = seq_len_interpolation_factor
        inv = 1.0 / (rotary_base ** (torch.arange(0, dim, 2, dtype=torch.float32) / dim))
        if rope_scaling:
            inv = _apply_llama3_scaling(inv, factor=rope_scaling_factor)
        self.register_buffer("inv_freq", inv, persistent=False)
        self._cache = {}

    def get_rotary_seq_len(self, inference_context, transformer, transformer_input, transformer_config,
                           packed_seq_params=None):
        if inference_context is not None:
            return inference_context.max_sequence_length
        if transformer is not None and getattr(transformer, "input_tensor", None) is not None:
            s = transformer.input_tensor.size(0)
        else:
            s = transformer_input.size(0)
        if transformer_config.sequence_parallel:
            s *= transformer_config.tensor_model_parallel_size
        return s * transformer_config.context_parallel_size

    def forward(self, max_seq_len: int, offset: int = 0, device=None) -> torch.Tensor:
        device = device or (torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else "cpu")
        key = (max_seq_len, offset, str(device))
        t = self._cache.get(key)
        if t is None:
            pos = torch.arange(max_seq_len, dtype=torch.float32, device=device) + offset
            if self.seq_len_interpolation_factor is not None:
                pos = pos / self.seq_len_interpolation_factor
            ang = torch.outer(pos, self.inv_freq.to(device))
            t = torch.cat([ang.cos(), ang.sin()], dim=-1).contiguous()
            if len(self._cache) > 8:
                self._cache.clear()
            self._cache[key] = t
        return t


class MultimodalRotaryEmbedding(nn.Module):
    """M-RoPE (Qwen2-VL): three position streams -- temporal, height, width -- each drive their own section of the rotary
    channels.  ``forward(position_ids [3, b, s], mrope_section)`` -> ``[s, b, 2 * dim/2]`` table of (cos | sin) with
    channel ``c`` taken from stream ``section_of(c) % 3``.  Parity: reference ``rotary_pos_embedding.py:224-310``."""

    def __init__(self, kv_channels: int, rotary_percent: float = 1.0, rotary_interleaved: bool = False,
                 seq_len_interpolation_factor: Optional[float] = None, rotary_base: float = 10000.0):
        super().__init__()
        dim = kv_channels if rotary_percent >= 1.0 else int(kv_channels * rotary_percent)
        self.dim, self.rotary_interleaved = dim, rotary_interleaved
        self.seq_len_interpolation_factor = seq_len_interpolation_factor
        self.register_buffer("inv_freq", 1.0 / (rotary_base ** (torch.arange(0, dim, 2, dtype=torch.float32) / dim)),
                             persistent=False)

    def forward(self, position_ids: torch.Tensor, mrope_section) -> torch.Tensor:
        if position_ids.dim() != 3 or position_ids.shape[0] != 3:
            raise ValueError("M-RoPE position ids must be [3, batch, seq] (temporal, height, width)")
        if sum(mrope_section) != self.dim // 2:
            raise ValueError(f"mrope_section {list(mrope_section)} must sum to rotary_dim / 2 = {self.dim // 2}")
        pos = position_ids.float()
        if self.seq_len_interpolation_factor is not None:
            pos = pos / self.seq_len_interpolation_factor
        ang = pos.unsqueeze(-1) * self.inv_freq.to(pos.device)                      # [3, b, s, dim/2]
        parts, c0 = [], 0
        for i, n in enumerate(mrope_section):
            parts.append(ang[i % 3, :, :, c0:c0 + n])
            c0 += n
        ang = torch.cat(parts, dim=-1).transpose(0, 1)                               # [s, b, dim/2]
        return torch.cat([ang.cos(), ang.sin()], dim=-1).contiguous()


class YarnRotaryEmbedding(RotaryEmbedding):
    """YaRN: NTK-by-parts interpolation of the rotary frequencies + attention temperature."""

    def __init__(self, kv_channels, rotary_percent=1.0, rotary_interleaved=False,
                 seq_len_interpolation_factor=None, rotary_base=10000.0, scaling_factor=1.0,
                 original_max_position_embeddings=4096, beta_fast=32.0, beta_slow=1.0, mscale=1.0,
                 mscale_all_dim=0.0, use_cpu_initialization=False):
        super().__init__(kv_channels, rotary_percent, rotary_interleaved, seq_len_interpolation_factor, rotary_base)
        dim = self.dim

        def corr_dim(num_rot):
            return (dim * math.log(original_max_position_embeddings / (num_rot * 2 * math.pi))) / (2 * math.log(rotary_base))

        low = max(math.floor(corr_dim(beta_fast)), 0)
        high = min(math.ceil(corr_dim(beta_slow)), dim - 1)
        ramp = ((torch.arange(dim // 2, dtype=torch.float32) - low) / max(high - low, 1e-3)).clamp(0, 1)
        extra = 1.0 / (rotary_base ** (torch.arange(0, dim, 2, dtype=torch.float32) / dim))
        inter = extra / scaling_factor
        mask = 1.0 - ramp
        self.register_buffer("inv_freq", inter * (1 - mask) + extra * mask, persistent=False)

        def get_mscale(scale, m):
            return 1.0 if scale <= 1 else 0.1 * m * math.log(scale) + 1.0

        self.concentration = get_mscale(scaling_factor, mscale) / get_mscale(scaling_factor, mscale_all_dim)

    def forward(self, max_seq_len, offset=0, device=None):
        t = super().forward(max_seq_len, offset, device)
        return t * self.concentration if self.concentration != 1.0 else t


class RelativePositionEmbedding(torch.nn.Module):
    """T5 bucketed relative position bias (reference ``models/common/embeddings/relative_pos_embedding.py:19``):
    ``forward(sq, sk) -> [1, heads, sq, sk]`` added to the attention scores.  Half of the buckets are exact
    small offsets, the rest are log-spaced up to ``max_distance``; the bidirectional (encoder) variant spends
    half of the buckets on each sign, the causal (decoder) variant only looks backwards."""

    def __init__(self, bidirectional: bool, init_method, num_attention_heads: int,
                 relative_attention_num_buckets: int = 32, relative_attention_max_distance: int = 128):
        super().__init__()
        self.bidirectional = bidirectional
        self.num_buckets, self.max_distance = relative_attention_num_buckets, relative_attention_max_distance
        self.relative_attention_bias = torch.nn.Embedding(self.num_buckets, num_attention_heads)
        init_method(self.relative_attention_bias.weight)

    def bucket(self, rel: torch.Tensor) -> torch.Tensor:
        """``rel = key_pos - query_pos`` (long) -> bucket id in ``[0, num_buckets)``."""
        n = self.num_buckets
        base = torch.zeros_like(rel)
        if self.bidirectional:
            n //= 2
            base = (rel > 0).long() * n
            dist = rel.abs()
        else:
            dist = (-rel).clamp_min(0)
        exact = n // 2
        log_b = exact + (torch.log(dist.clamp_min(1).float() / exact) / math.log(self.max_distance / exact)
                         * (n - exact)).long()
        return base + torch.where(dist < exact, dist, log_b.clamp_max(n - 1))

    def forward(self, query_seq_length: int, key_seq_length: int) -> torch.Tensor:
        dev = self.relative_attention_bias.weight.device
        rel = torch.arange(key_seq_length, device=dev)[None, :] - torch.arange(query_seq_length, device=dev)[:, None]
        return self.relative_attention_bias(self.bucket(rel)).permute(2, 0, 1).unsqueeze(0)

    @staticmethod
    def get_relative_seq_len(inference_context, transformer, transformer_input, transformer_config) -> int:
        if inference_context is not None:
            return inference_context.max_sequence_length
        t = getattr(transformer, "input_tensor", None)
        s = (t if t is not None else transformer_input).size(0)
        return s * transformer_config.tensor_model_parallel_size if transformer_config.sequence_parallel else s


# -- function forms of the reference's ``rope_utils.py`` ---------------------------------------------------------------------------
def _rotate_half(x: torch.Tensor, rotary_interleaved: bool) -> torch.Tensor:
    if not rotary_interleaved:
        x1, x2 = torch.chunk(x, 2, dim=-1)
        return torch.cat((-x2, x1), dim=-1)
    x1, x2 = x[..., ::2], x[..., 1::2]
    return torch.stack((-x2, x1), dim=-1).flatten(-2)


def apply_rotary_pos_emb_with_cos_sin(t: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor, rotary_interleaved: bool = False):
    """``t [..., d]`` rotated by per-position ``cos`` / ``sin [..., rot]`` (``rot <= d``; the remaining channels pass through)."""
    rot = cos.shape[-1]
    tr, tp = t[..., :rot], t[..., rot:]
    out = tr.float() * cos.float() + _rotate_half(tr.float(), rotary_interleaved) * sin.float()
    return torch.cat((out.to(t.dtype), tp), dim=-1)


def apply_rotary_pos_emb(t: torch.Tensor, freqs: torch.Tensor, config=None, cu_seqlens=None, mscale: float = 1.0, cp_group=None):
    """``t [s, b, h, d]`` with either Megatron's ``freqs [s, 1, 1, rot]`` (the angles, each twice) or this framework's
    ``[s, rot]`` cos | sin table (what ``RotaryEmbedding.forward`` returns here).  On a GPU the RoPE kernel is used; packed
    sequences (``cu_seqlens``) apply each sequence's own positions."""
    interleaved = bool(getattr(config, "rotary_interleaved", False))
    if cu_seqlens is not None:                      # THD: t [t, h, d]; positions restart at every sequence boundary
        bounds = [int(x) for x in (cu_seqlens.tolist() if torch.is_tensor(cu_seqlens) else cu_seqlens)]
        parts = [apply_rotary_pos_emb(t[a:b].unsqueeze(1), freqs[: b - a], config, None, mscale).squeeze(1)
                 for a, b in zip(bounds[:-1], bounds[1:])]
        return torch.cat(parts, dim=0)
    if freqs.dim() == 2:
        # this framework's own table (``RotaryEmbedding.forward``): ``[s, rot]`` holding cos | sin of the rot / 2 angles
        from ... import ops

        out = ops.rope(t, freqs, interleaved) if t.dim() == 4 else ops.rope(t.unsqueeze(1), freqs, interleaved).squeeze(1)
        return out if mscale == 1.0 else out * mscale
    if t.is_cuda and mscale == 1.0 and t.dim() == 4:
        from ...core_api.te import fused_apply_rotary_pos_emb

        return fused_apply_rotary_pos_emb(t, freqs, interleaved=interleaved)
    f = freqs.float()
    return apply_rotary_pos_emb_with_cos_sin(t, (f.cos() * mscale).to(t.dtype), (f.sin() * mscale).to(t.dtype), interleaved)


def get_pos_emb_on_this_cp_rank(pos_emb: torch.Tensor, seq_dim: int, cp_group=None) -> torch.Tensor:
    """This context-parallel rank's two zig-zag chunks (``rank`` and ``2 cp - rank - 1``) of a position-embedding table."""
    from ...parallel import state as ps

    cp, rank = ps.get_context_parallel_world_size(), ps.get_context_parallel_rank()
    if cp == 1:
        return pos_emb
    idx = torch.tensor([rank, 2 * cp - rank - 1], device=pos_emb.device)
    shape = pos_emb.shape
    x = pos_emb.view(*shape[:seq_dim], 2 * cp, -1, *shape[seq_dim + 1:]).index_select(seq_dim, idx)
    return x.view(*shape[:seq_dim], -1, *shape[seq_dim + 1:])
