"""Multi-latent attention (DeepSeek-V2/V3): low-rank Q / KV projections with a decoupled rotary
sub-head that is shared across heads.

Parity: reference ``transformer/multi_latent_attention.py`` (``MultiLatentAttention :44``,
``MLASelfAttention :228``: ``linear_q_proj`` or ``linear_q_down_proj -> q_layernorm ->
linear_q_up_proj``; ``linear_kv_down_proj`` producing ``kv_lora_rank + qk_pos_emb_head_dim``
channels; ``kv_layernorm``; ``linear_kv_up_proj`` producing per-head ``qk_head_dim + v_head_dim``;
YaRN rotary embedding with the mscale softmax temperature; ``linear_proj`` from
``heads * v_head_dim``) and ``MLATransformerConfig`` (``transformer_config.py:1072``).

The query/key head width (``qk_head_dim + qk_pos_emb_head_dim``) differs from the value width;
V is zero-padded to the QK width for the fused kernel and the context is sliced back.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional, Union

import torch

from .. import ops
from ..models.common.embeddings import RotaryEmbedding, YarnRotaryEmbedding
from ..parallel import state as ps
from ..scan.tracer import scope as _scope
from .module import MegatronModule
from .spec_utils import ModuleSpec, build_module


@dataclass
class MLASelfAttentionSubmodules:
    linear_q_proj: Union[ModuleSpec, type] = None
    linear_q_down_proj: Union[ModuleSpec, type] = None
    linear_q_up_proj: Union[ModuleSpec, type] = None
    linear_kv_down_proj: Union[ModuleSpec, type] = None
    linear_kv_up_proj: Union[ModuleSpec, type] = None
    core_attention: Union[ModuleSpec, type] = None
    linear_proj: Union[ModuleSpec, type] = None
    q_layernorm: Union[ModuleSpec, type] = None
    kv_layernorm: Union[ModuleSpec, type] = None


def _yarn_mscale(scale: float, mscale: float) -> float:
    return 1.0 if scale <= 1 else 0.1 * mscale * math.log(scale) + 1.0


class MLASelfAttention(MegatronModule):
    def __init__(self, config, submodules: MLASelfAttentionSubmodules, layer_number: int, attn_mask_type="causal",
                 cp_comm_type: Optional[str] = None):
        super().__init__(config)
        self.layer_number = layer_number
        self.attn_mask_type = attn_mask_type
        tp = ps.get_tensor_model_parallel_world_size()
        self.nh = config.num_attention_heads
        self.nh_local = self.nh // tp
        self.qk_dim, self.rope_dim, self.v_dim = config.qk_head_dim, config.qk_pos_emb_head_dim, config.v_head_dim
        self.q_head_dim = self.qk_dim + self.rope_dim
        self.kv_lora_rank, self.q_lora_rank = config.kv_lora_rank, config.q_lora_rank
        ms = _yarn_mscale(config.rotary_scaling_factor, config.mscale) if config.rope_type == "yarn" else 1.0
        self.softmax_scale = ms * ms / math.sqrt(self.q_head_dim)
        if config.rope_type == "yarn":
            self.rotary_pos_emb = YarnRotaryEmbedding(
                self.rope_dim, rotary_base=config.rotary_base, scaling_factor=config.rotary_scaling_factor,
                original_max_position_embeddings=config.max_position_embeddings, beta_fast=config.beta_fast,
                beta_slow=config.beta_slow, mscale=config.mscale, mscale_all_dim=config.mscale_all_dim)
        else:
            self.rotary_pos_emb = RotaryEmbedding(self.rope_dim, rotary_base=config.rotary_base)
        col = dict(config=config, init_method=config.init_method, bias=False, skip_bias_add=False, is_expert=False)
        if self.q_lora_rank is None:
            self.linear_q_proj = build_module(submodules.linear_q_proj, config.hidden_size, self.nh * self.q_head_dim,
                                              gather_output=False, **col)
        else:
            self.linear_q_down_proj = build_module(submodules.linear_q_down_proj, config.hidden_size, self.q_lora_rank,
                                                   gather_output=True, **col)
            self.linear_q_up_proj = build_module(submodules.linear_q_up_proj, self.q_lora_rank, self.nh * self.q_head_dim,
                                                 gather_output=False, **col)
            self.q_layernorm = build_module(submodules.q_layernorm, config=config, hidden_size=self.q_lora_rank,
                                            eps=config.layernorm_epsilon)
        self.linear_kv_down_proj = build_module(submodules.linear_kv_down_proj, config.hidden_size,
                                                self.kv_lora_rank + self.rope_dim, gather_output=True, **col)
        self.linear_kv_up_proj = build_module(submodules.linear_kv_up_proj, self.kv_lora_rank,
                                              self.nh * (self.qk_dim + self.v_dim), gather_output=False,
                                              **col)
        self.kv_layernorm = build_module(submodules.kv_layernorm, config=config, hidden_size=self.kv_lora_rank,
                                         eps=config.layernorm_epsilon)
        # the up-projections and the latent norms see the full (gathered, TP-replicated) sequence:
        # no sequence-parallel collectives there, dgrad is all-reduced, norm grads are not TP-summed
        for m in (getattr(self, "linear_q_up_proj", None), self.linear_kv_up_proj):
            if m is not None:
                m.sequence_parallel = False
                m.allreduce_dgrad = tp > 1
        for m in (getattr(self, "q_layernorm", None), self.kv_layernorm):
            for prm in (m.parameters() if m is not None else ()):
                setattr(prm, "sequence_parallel", False)
        self.linear_proj = build_module(submodules.linear_proj, self.nh * self.v_dim, config.hidden_size, config=config,
                                        init_method=config.output_layer_init_method, bias=config.add_bias_linear,
                                        input_is_parallel=True, skip_bias_add=True, is_expert=False)
        self.dropout_p = config.attention_dropout

    def _qkv(self, hidden_states, inference_context=None):
        if self.q_lora_rank is None:
            q, _ = self.linear_q_proj(hidden_states)
        else:
            qc, _ = self.linear_q_down_proj(hidden_states)        # [s, b, r_q] (sequence gathered under SP)
            q, _ = self.linear_q_up_proj(self.q_layernorm(qc))
        s, b = q.shape[0], q.shape[1]
        q = q.view(s, b, self.nh_local, self.q_head_dim)
        kvc, _ = self.linear_kv_down_proj(hidden_states)           # [s, b, r_kv + rope]
        kv_c, k_pe = torch.split(kvc, [self.kv_lora_rank, self.rope_dim], dim=-1)
        kv, _ = self.linear_kv_up_proj(self.kv_layernorm(kv_c.contiguous()))
        kv = kv.view(kv.shape[0], b, self.nh_local, self.qk_dim + self.v_dim)
        k_nope, v = torch.split(kv, [self.qk_dim, self.v_dim], dim=-1)
        q_nope, q_pe = torch.split(q, [self.qk_dim, self.rope_dim], dim=-1)
        off = inference_context.sequence_len_offset if inference_context is not None else 0
        cs = self.rotary_pos_emb(off + kv.shape[0], device=q.device)
        q_pe = ops.rope(q_pe.contiguous(), cs[off:off + s].contiguous(), self.config.rotary_interleaved)
        k_pe = ops.rope(k_pe.reshape(kv.shape[0], b, 1, self.rope_dim).contiguous(), cs[off:off + kv.shape[0]].contiguous(),
                        self.config.rotary_interleaved)
        q = torch.cat([q_nope, q_pe], dim=-1)
        k = torch.cat([k_nope, k_pe.expand(-1, -1, self.nh_local, -1)], dim=-1)
        return q, k, v

    def forward(self, hidden_states, attention_mask=None, key_value_states=None, inference_context=None,
                rotary_pos_emb=None, rotary_pos_cos=None, rotary_pos_sin=None, attention_bias=None,
                packed_seq_params=None, sequence_len_offset=None, inference_params=None):
        inference_context = inference_context or inference_params
        with _scope("attention"):
            q, k, v = self._qkv(hidden_states, inference_context)
            sq, b = q.shape[0], q.shape[1]
            causal = str(self.attn_mask_type).endswith("causal")
            if inference_context is not None:
                cache = inference_context.key_value_memory_dict
                if self.layer_number not in cache:
                    ms, mb = inference_context.max_sequence_length, inference_context.max_batch_size
                    cache[self.layer_number] = (k.new_empty(ms, mb, *k.shape[2:]), v.new_empty(ms, mb, *v.shape[2:]))
                kc, vc = cache[self.layer_number]
                s0 = inference_context.sequence_len_offset
                kc[s0:s0 + sq, :b], vc[s0:s0 + sq, :b] = k, v
                k, v = kc[:s0 + sq, :b], vc[:s0 + sq, :b]
                if s0 > 0 and sq == 1:
                    causal = False
            pad = self.q_head_dim - self.v_dim
            vp = torch.nn.functional.pad(v, (0, pad)) if pad > 0 else v
            o = ops.attention(q.permute(1, 2, 0, 3), k.permute(1, 2, 0, 3), vp.permute(1, 2, 0, 3), self.softmax_scale,
                              causal, self.dropout_p if self.training else 0.0, self.training,
                              None if causal else attention_mask)
            ctx = o.permute(2, 0, 1, 3)[..., :self.v_dim].reshape(sq, b, self.nh_local * self.v_dim)
            out, bias = self.linear_proj(ctx)
        return out, bias
