"""Self / cross attention with GQA, RoPE, KV cache and a fused core.

Parity: reference ``transformer/attention.py`` (``Attention`` :471-682, ``SelfAttention`` :845
incl. the packed QKV layout ``[s, b, ng, (np/ng + 2) * hn]`` of
``get_query_key_value_tensors`` :973, ``CrossAttention`` :1029) and
``dot_product_attention.py:106-208`` -- but the core is one flash-attention kernel
(``ops.attention``) instead of baddbmm -> softmax -> bmm, so the MegaScan / MegaScope add-ons
are not tied to an O(s^2) "local" path.  MegaScope probability export (FlagType 2) switches a
layer to the materialising reference path only while that flag is set.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional, Union

import torch
import torch.nn as nn

from ..parallel import state as ps
from ..parallel.random import get_cuda_rng_tracker
from .. import ops
from ..scan.tracer import scope as _scope
from ..scope import hooks as scope_hooks
from .module import MegatronModule
from .spec_utils import ModuleSpec, build_module


@dataclass
class SelfAttentionSubmodules:
    linear_qkv: Union[ModuleSpec, type] = None
    core_attention: Union[ModuleSpec, type] = None
    linear_proj: Union[ModuleSpec, type] = None
    q_layernorm: Union[ModuleSpec, type] = None
    k_layernorm: Union[ModuleSpec, type] = None


@dataclass
class CrossAttentionSubmodules:
    linear_q: Union[ModuleSpec, type] = None
    linear_kv: Union[ModuleSpec, type] = None
    core_attention: Union[ModuleSpec, type] = None
    linear_proj: Union[ModuleSpec, type] = None


def _packed_positions(cu_seqlens: torch.Tensor, total: int, device) -> torch.Tensor:
    """``[0..len_0-1, 0..len_1-1, ...]`` (padding after the last sequence keeps counting) for THD rotary embeddings."""
    cu = cu_seqlens.to(device=device, dtype=torch.long)
    idx = torch.arange(total, device=device)
    seg = torch.bucketize(idx, cu[1:], right=True).clamp_max(cu.numel() - 2)
    return idx - cu[seg]


class DotProductAttention(MegatronModule):
    """Core attention on ``[s, b, np, hn]`` tensors -> ``[s, b, np*hn]``."""

    def __init__(self, config, layer_number: int, attn_mask_type="causal", attention_type="self",
                 attention_dropout: Optional[float] = None, softmax_scale: Optional[float] = None,
                 cp_comm_type: Optional[str] = None):
        super().__init__(config)
        self.layer_number = max(1, layer_number)
        self.attn_mask_type = attn_mask_type
        self.attention_type = attention_type
        tp = ps.get_tensor_model_parallel_world_size()
        self.hidden_size_per_partition = config.kv_channels * config.num_attention_heads // tp
        self.hn = config.kv_channels
        self.np = config.num_attention_heads // tp
        self.ng = config.num_query_groups // tp
        self.softmax_scale = softmax_scale if softmax_scale is not None else (
            config.softmax_scale if config.softmax_scale is not None else 1.0 / math.sqrt(self.hn))
        if config.apply_query_key_layer_scaling:
            # the extra 1/layer_number coefficient cancels inside the fp32 softmax
            pass
        self.dropout_p = config.attention_dropout if attention_dropout is None else attention_dropout
        cct = cp_comm_type if cp_comm_type is not None else config.cp_comm_type
        if isinstance(cct, (list, tuple)):
            cct = cct[self.layer_number - 1]
        self.cp_comm_type = cct or "p2p"
        self.cp_size = config.context_parallel_size

    def _attend(self, q, k, v, causal, dropout_p, mask, bias=None):
        if bias is not None:
            if bias.shape[1] != q.shape[1]:        # full-head bias under TP: this rank's head slice
                r = ps.get_tensor_model_parallel_rank()
                bias = bias[:, r * q.shape[1]:(r + 1) * q.shape[1]]
            return ops.attention(q, k, v, self.softmax_scale, causal, dropout_p, self.training, mask, bias=bias)
        if self.cp_size > 1 and self.attention_type == "self":
            from ..parallel.context_parallel import cp_attention

            return cp_attention(q, k, v, self.softmax_scale, causal, self.cp_comm_type, dropout_p, self.training)
        return ops.attention(q, k, v, self.softmax_scale, causal, dropout_p, self.training, mask)

    def forward(self, query, key, value, attention_mask=None, attn_mask_type=None, attention_bias=None,
                packed_seq_params=None):
        if packed_seq_params is not None:          # THD: [t, 1, np, hn] with sequence boundaries in cu_seqlens
            if query.shape[1] != 1:
                raise ValueError("packed sequences expect micro-batch 1 (all sequences concatenated along tokens)")
            causal = str(attn_mask_type or self.attn_mask_type).endswith("causal")
            host = packed_seq_params.boundaries() if hasattr(packed_seq_params, "boundaries") else None
            o = ops.attention_varlen(query[:, 0], key[:, 0], value[:, 0], host or packed_seq_params.cu_seqlens_q,
                                     None if host else packed_seq_params.cu_seqlens_kv, self.softmax_scale, causal,
                                     self.dropout_p if self.training else 0.0, self.training)
            return o.reshape(query.shape[0], 1, self.np * o.shape[-1])
        mask_type = attn_mask_type or self.attn_mask_type
        causal = str(mask_type).endswith("causal")
        sq, b = query.shape[0], query.shape[1]
        q = query.permute(1, 2, 0, 3)  # [b, np, sq, hn]
        k = key.permute(1, 2, 0, 3)
        v = value.permute(1, 2, 0, 3)
        want_probs = scope_hooks.wants(2, self.layer_number)
        mask = None if causal else attention_mask
        if want_probs:
            o, probs = ops.attention_ref(q, k, v, self.softmax_scale, causal, 0.0, return_probs=True, mask=mask,
                                         bias=attention_bias)
            scope_hooks.tik_tensor(2, self.layer_number, probs)
        else:
            if self.dropout_p > 0.0 and self.training:
                with get_cuda_rng_tracker().fork() if "model-parallel-rng" in get_cuda_rng_tracker().states_ \
                        else torch.random.fork_rng(devices=[]):
                    o = self._attend(q, k, v, causal, self.dropout_p, mask, attention_bias)
            else:
                o = self._attend(q, k, v, causal, 0.0, mask, attention_bias)
        ctx = o.permute(2, 0, 1, 3).reshape(sq, b, self.np * self.hn)
        scope_hooks.tik_tensor(3, self.layer_number, ctx)
        return ctx


class Attention(MegatronModule):
    def _packed_self_attention(self, *a, **k):
        return None

    def __init__(self, config, submodules, layer_number: int, attn_mask_type, attention_type: str,
                 cp_comm_type: Optional[str] = None):
        super().__init__(config)
        self.layer_number = layer_number
        self.attn_mask_type = attn_mask_type
        self.attention_type = attention_type
        self.query_projection_size = config.kv_channels * config.num_attention_heads
        self.kv_projection_size = config.kv_channels * config.num_query_groups
        tp = ps.get_tensor_model_parallel_world_size()
        self.hidden_size_per_attention_head = config.kv_channels
        self.num_attention_heads_per_partition = config.num_attention_heads // tp
        self.num_query_groups_per_partition = config.num_query_groups // tp
        self.core_attention = build_module(
            submodules.core_attention, config=config, layer_number=layer_number,
            attn_mask_type=attn_mask_type, attention_type=attention_type)
        self.linear_proj = build_module(
            submodules.linear_proj, self.query_projection_size, config.hidden_size, config=config,
            init_method=config.output_layer_init_method, bias=config.add_bias_linear,
            input_is_parallel=True, skip_bias_add=True, is_expert=False, tp_comm_buffer_name="proj")
        self.checkpoint_core_attention = (config.recompute_granularity == "selective"
                                          and "core_attn" in (config.recompute_modules or []))

    # -- KV cache (static inference) ------------------------------------------
    def _adjust_kv_for_inference(self, inference_context, key, value):
        if inference_context is None:
            return key, value
        cache = inference_context.key_value_memory_dict
        if self.layer_number not in cache:
            max_s, max_b = inference_context.max_sequence_length, inference_context.max_batch_size
            shape = (max_s, max_b, key.shape[2], key.shape[3])
            cache[self.layer_number] = (torch.empty(shape, dtype=key.dtype, device=key.device),
                                        torch.empty(shape, dtype=value.dtype, device=value.device))
        kc, vc = cache[self.layer_number]
        b0 = inference_context.batch_size_offset
        b1 = b0 + key.shape[1]
        s0 = inference_context.sequence_len_offset
        s1 = s0 + key.shape[0]
        kc[s0:s1, b0:b1] = key
        vc[s0:s1, b0:b1] = value
        return kc[:s1, b0:b1], vc[:s1, b0:b1]

    def forward(self, hidden_states, attention_mask=None, key_value_states=None, inference_context=None,
                rotary_pos_emb=None, rotary_pos_cos=None, rotary_pos_sin=None, attention_bias=None,
                packed_seq_params=None, sequence_len_offset=None, inference_params=None):
        inference_context = inference_context or inference_params
        with _scope("attention"):
            packed = None if attention_bias is not None else self._packed_self_attention(
                hidden_states, attention_mask, rotary_pos_emb, inference_context, packed_seq_params)
            if packed is not None:
                output, bias = self.linear_proj(packed)
                return output, bias
            query, key, value = self.get_query_key_value_tensors(hidden_states, key_value_states)
            if inference_context is not None and not inference_context.is_static_batching():
                return self._dynamic_batch_attention(query, key, value, rotary_pos_emb, inference_context)
            if rotary_pos_emb is not None:
                q_cs, k_cs = rotary_pos_emb if isinstance(rotary_pos_emb, tuple) else (rotary_pos_emb, rotary_pos_emb)
                if packed_seq_params is not None:      # positions restart at every packed sequence
                    q_cs = q_cs[_packed_positions(packed_seq_params.cu_seqlens_q, query.shape[0], q_cs.device)]
                    k_cs = k_cs[_packed_positions(packed_seq_params.cu_seqlens_kv if packed_seq_params.cu_seqlens_kv is not None
                                                  else packed_seq_params.cu_seqlens_q, key.shape[0], k_cs.device)]
                elif inference_context is not None:
                    off = inference_context.sequence_len_offset
                    q_cs = q_cs[off:off + query.shape[0]]
                    k_cs = k_cs[off:off + key.shape[0]]
                query = ops.rope(query, q_cs.contiguous(), self.config.rotary_interleaved)
                key = ops.rope(key, k_cs.contiguous(), self.config.rotary_interleaved)
            key, value = self._adjust_kv_for_inference(inference_context, key, value)
            mask_type = self.attn_mask_type
            if inference_context is not None and inference_context.sequence_len_offset > 0 and query.shape[0] == 1:
                mask_type = "no_mask"  # single-token decode attends to the whole cache
            if self.checkpoint_core_attention and self.training:
                from ..parallel.random import checkpoint

                core = checkpoint(lambda q, k, v: self.core_attention(q, k, v, attention_mask, mask_type,
                                                                      attention_bias=attention_bias),
                                  False, query, key, value)
            else:
                core = self.core_attention(query, key, value, attention_mask, attn_mask_type=mask_type,
                                           attention_bias=attention_bias, packed_seq_params=packed_seq_params)
            output, bias = self.linear_proj(core)
        return output, bias


    def _dynamic_batch_attention(self, query, key, value, rotary_pos_emb, ctx):
        """Packed token batch ``[T, 1, ...]`` against the paged KV cache (``inference/dynamic.py``)."""
        if rotary_pos_emb is not None:
            q_cs, k_cs = rotary_pos_emb if isinstance(rotary_pos_emb, tuple) else (rotary_pos_emb, rotary_pos_emb)
            query = ops.rope(query, q_cs[ctx.positions].contiguous(), self.config.rotary_interleaved)
            key = ops.rope(key, k_cs[ctx.positions].contiguous(), self.config.rotary_interleaved)
        o = ctx.attend(self.layer_number, query[:, 0], key[:, 0], value[:, 0], self.core_attention.softmax_scale)
        return self.linear_proj(o.reshape(o.shape[0], 1, -1))


class SelfAttention(Attention):
    def _packed_self_attention(self, hidden_states, attention_mask, rotary_pos_emb, inference_context, packed_seq_params):
        """MHA training fast path: attention directly on the packed QKV projection (no split / concat)."""
        core = self.core_attention
        if (self.num_query_groups_per_partition != self.num_attention_heads_per_partition or rotary_pos_emb is not None
                or inference_context is not None or packed_seq_params is not None or self.q_layernorm is not None
                or not isinstance(core, DotProductAttention) or core.cp_size > 1 or self.checkpoint_core_attention
                or not str(self.attn_mask_type).endswith("causal") or (core.dropout_p > 0.0 and self.training)
                or scope_hooks.wants(1, self.layer_number) or scope_hooks.wants(2, self.layer_number)
                or scope_hooks.wants(3, self.layer_number) or not hidden_states.is_cuda):
            return None
        mixed, _ = self.linear_qkv(hidden_states)
        s_, b_ = mixed.shape[0], mixed.shape[1]
        np_, hn = self.num_attention_heads_per_partition, self.hidden_size_per_attention_head
        o = ops.attention_packed_qkv(mixed.view(s_, b_, np_, 3 * hn), core.softmax_scale, True)
        if o is None:
            self._mixed_cache = mixed        # reuse the projection in the generic path
            return None
        return o.permute(2, 0, 1, 3).reshape(s_, b_, np_ * hn)


    def __init__(self, config, submodules: SelfAttentionSubmodules, layer_number: int,
                 attn_mask_type="causal", cp_comm_type=None):
        super().__init__(config, submodules, layer_number, attn_mask_type, "self", cp_comm_type)
        self.linear_qkv = build_module(
            submodules.linear_qkv, config.hidden_size, self.query_projection_size + 2 * self.kv_projection_size,
            config=config, init_method=config.init_method, gather_output=False,
            bias=config.add_bias_linear or config.add_qkv_bias, skip_bias_add=False, is_expert=False,
            tp_comm_buffer_name="qkv")
        if submodules.q_layernorm is not None:
            self.q_layernorm = build_module(submodules.q_layernorm, hidden_size=self.hidden_size_per_attention_head,
                                            config=config, eps=config.layernorm_epsilon)
            self.k_layernorm = build_module(submodules.k_layernorm, hidden_size=self.hidden_size_per_attention_head,
                                            config=config, eps=config.layernorm_epsilon)
        else:
            self.q_layernorm = self.k_layernorm = None

    def get_query_key_value_tensors(self, hidden_states, key_value_states=None):
        mixed = getattr(self, "_mixed_cache", None)
        self._mixed_cache = None
        if mixed is None:
            mixed, _ = self.linear_qkv(hidden_states)  # [s, b, ng*(np/ng+2)*hn]
        scope_hooks.tik_tensor(1, self.layer_number, mixed)
        ng = self.num_query_groups_per_partition
        hn = self.hidden_size_per_attention_head
        qpg = self.num_attention_heads_per_partition // ng
        mixed = mixed.view(mixed.shape[0], mixed.shape[1], ng, (qpg + 2) * hn)
        q, k, v = torch.split(mixed, [qpg * hn, hn, hn], dim=3)
        q = q.reshape(q.shape[0], q.shape[1], -1, hn)
        if self.q_layernorm is not None:
            q = self.q_layernorm(q.contiguous())
            k = self.k_layernorm(k.contiguous())
        return q, k, v


class CrossAttention(Attention):
    def __init__(self, config, submodules: CrossAttentionSubmodules, layer_number: int,
                 attn_mask_type="padding", cp_comm_type=None):
        super().__init__(config, submodules, layer_number, attn_mask_type, "cross", cp_comm_type)
        if config.num_query_groups != config.num_attention_heads:
            raise ValueError("group query attention is not supported in cross attention")
        self.linear_q = build_module(submodules.linear_q, config.hidden_size, self.query_projection_size,
                                     config=config, init_method=config.init_method, gather_output=False,
                                     bias=config.add_bias_linear, skip_bias_add=False, is_expert=False)
        self.linear_kv = build_module(submodules.linear_kv, config.hidden_size, 2 * self.kv_projection_size,
                                      config=config, init_method=config.init_method, gather_output=False,
                                      bias=config.add_bias_linear, skip_bias_add=False, is_expert=False)

    def get_query_key_value_tensors(self, hidden_states, key_value_states):
        kv, _ = self.linear_kv(key_value_states)
        hn = self.hidden_size_per_attention_head
        kv = kv.view(kv.shape[0], kv.shape[1], self.num_attention_heads_per_partition, 2 * hn)
        k, v = torch.split(kv, [hn, hn], dim=3)
        q, _ = self.linear_q(hidden_states)
        q = q.view(q.shape[0], q.shape[1], self.num_attention_heads_per_partition, hn)
        return q, k, v
