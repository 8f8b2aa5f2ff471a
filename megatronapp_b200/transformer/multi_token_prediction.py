"""Multi-token prediction (DeepSeek-V3 style auxiliary heads).

Parity: reference ``transformer/multi_token_prediction.py`` -- ``MultiTokenPredictionLayer :316``
(RMSNorm of the previous depth's hidden state and of the embedding of the next token, concatenate,
project ``2h -> h``, one transformer layer, final norm), ``MultiTokenPredictionBlock :580`` (depth
loop sharing the embedding and the output head with the main model, per-depth loss on labels
rolled left, scaled by ``mtp_loss_scaling_factor / mtp_num_layers``), ``MTPLossAutoScaler``
(attaches the auxiliary loss to the main hidden states so it needs no extra plumbing through the
pipeline schedule) and ``MTPLossLoggingHelper``.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Union

import torch

from ..parallel import state as ps
from ..parallel.mappings import gather_from_sequence_parallel_region, scatter_to_sequence_parallel_region
from .module import MegatronModule
from .spec_utils import ModuleSpec, build_module


def roll_tensor(t: torch.Tensor, shifts: int = -1, dims: int = -1):
    """Roll left and zero the vacated tail; returns (rolled, sum) like the reference helper."""
    r = torch.roll(t, shifts=shifts, dims=dims)
    idx = [slice(None)] * r.dim()
    idx[dims] = slice(shifts, None) if shifts < 0 else slice(0, shifts)
    r[tuple(idx)] = 0
    return r, r.sum()


class MTPLossAutoScaler(torch.autograd.Function):
    """Identity on the main activations whose backward also seeds d(mtp_loss) = scale."""

    main_loss_backward_scale: torch.Tensor = torch.tensor(1.0)

    @staticmethod
    def forward(ctx, output: torch.Tensor, mtp_loss: torch.Tensor):
        ctx.save_for_backward(mtp_loss)
        return output

    @staticmethod
    def backward(ctx, grad_output: torch.Tensor):
        (mtp_loss,) = ctx.saved_tensors
        scale = MTPLossAutoScaler.main_loss_backward_scale
        return grad_output, torch.ones_like(mtp_loss) * scale.to(mtp_loss.device).reshape(-1)[0]

    @staticmethod
    def set_loss_scale(scale: torch.Tensor):
        MTPLossAutoScaler.main_loss_backward_scale = scale


class MTPLossLoggingHelper:
    tracker = {}

    @staticmethod
    def save_loss_to_tracker(loss: torch.Tensor, layer_number: int, num_layers: int, reduce_group=None, avg_group=None):
        tr = MTPLossLoggingHelper.tracker
        if "values" not in tr:
            tr["values"] = torch.zeros(num_layers, device=loss.device)
        tr["values"][layer_number] += loss.detach()
        tr["reduce_group"], tr["avg_group"] = reduce_group, avg_group

    @staticmethod
    def clean_loss_in_tracker():
        if "values" in MTPLossLoggingHelper.tracker:
            MTPLossLoggingHelper.tracker["values"].zero_()

    @staticmethod
    def track_mtp_metrics(loss_scale, iteration, writer=None, total_loss_dict=None):
        tr = MTPLossLoggingHelper.tracker
        if "values" not in tr:
            return
        vals = tr["values"].float() * loss_scale
        for i, v in enumerate(vals.tolist()):
            name = f"mtp_{i + 1} loss"
            if total_loss_dict is not None:
                total_loss_dict[name] = total_loss_dict.get(name, 0.0) + v
            if writer is not None:
                writer.add_scalar(name, v, iteration)
        MTPLossLoggingHelper.clean_loss_in_tracker()


@dataclass
class MultiTokenPredictionLayerSubmodules:
    enorm: Union[ModuleSpec, type] = None
    hnorm: Union[ModuleSpec, type] = None
    eh_proj: Union[ModuleSpec, type] = None
    transformer_layer: Union[ModuleSpec, type] = None
    layer_norm: Union[ModuleSpec, type] = None


class MultiTokenPredictionLayer(MegatronModule):
    def __init__(self, config, submodules: MultiTokenPredictionLayerSubmodules, layer_number: int = 1):
        super().__init__(config)
        self.layer_number = layer_number
        eps = config.layernorm_epsilon
        self.enorm = build_module(submodules.enorm, config=config, hidden_size=config.hidden_size, eps=eps)
        self.hnorm = build_module(submodules.hnorm, config=config, hidden_size=config.hidden_size, eps=eps)
        # the concatenated [e ; h] input is not sequence-sharded consistently with a fused AG, so the
        # projection runs on the gathered sequence and re-scatters
        self.eh_proj = build_module(submodules.eh_proj, 2 * config.hidden_size, config.hidden_size, config=config,
                                    init_method=config.init_method, gather_output=False, bias=False,
                                    skip_bias_add=False, is_expert=False)
        self.transformer_layer = build_module(submodules.transformer_layer, config=config)
        self.final_layernorm = build_module(submodules.layer_norm, config=config, hidden_size=config.hidden_size, eps=eps)

    def forward(self, decoder_input, hidden_states, attention_mask=None, rotary_pos_emb=None, **kw):
        e = self.enorm(decoder_input)
        h = self.hnorm(hidden_states)
        x = torch.cat((e, h), dim=-1)
        x, _ = self.eh_proj(x)
        tp = ps.get_tensor_model_parallel_world_size()
        if tp > 1:   # column-parallel output -> full hidden
            from ..parallel.mappings import all_gather_last_dim_from_tensor_parallel_region

            x = all_gather_last_dim_from_tensor_parallel_region(x)
            if self.config.sequence_parallel:   # the column linear gathered the sequence; shard it again
                x = scatter_to_sequence_parallel_region(x)
        x, _ = self.transformer_layer(x, attention_mask=attention_mask, rotary_pos_emb=rotary_pos_emb)
        return self.final_layernorm(x)


@dataclass
class MultiTokenPredictionBlockSubmodules:
    layer_specs: list = None


class MultiTokenPredictionBlock(MegatronModule):
    def __init__(self, config, spec: Union[ModuleSpec, MultiTokenPredictionBlockSubmodules]):
        super().__init__(config)
        sub = spec.submodules if isinstance(spec, ModuleSpec) else spec
        self.layers = torch.nn.ModuleList(
            [build_module(ls, config=config, layer_number=i + 1) for i, ls in enumerate(sub.layer_specs)])
        self.mtp_loss_scaling_factor = config.mtp_loss_scaling_factor if config.mtp_loss_scaling_factor is not None else 0.1

    def forward(self, input_ids, position_ids, hidden_states, attention_mask, labels=None, loss_mask=None,
                rotary_pos_emb=None, embedding=None, output_layer=None, output_weight=None, compute_loss=None, **kw):
        """Returns the main hidden states with the (scaled) MTP loss attached to their graph."""
        assert labels is not None, "MTP needs labels (it is a training-only auxiliary objective)"
        if loss_mask is None:
            loss_mask = torch.ones_like(labels, dtype=torch.float32)
        main_hidden = hidden_states
        n = len(self.layers)
        for i, layer in enumerate(self.layers):
            input_ids, _ = roll_tensor(input_ids, -1, -1)
            labels, _ = roll_tensor(labels, -1, -1)
            loss_mask, ntok = roll_tensor(loss_mask, -1, -1)
            dec_in = embedding(input_ids=input_ids, position_ids=position_ids)
            hidden_states = layer(dec_in, hidden_states, attention_mask=attention_mask, rotary_pos_emb=rotary_pos_emb)
            logits, _ = output_layer(hidden_states, weight=output_weight)
            loss = compute_loss(labels, logits)                        # [b, s]
            loss = (loss * loss_mask).sum() / ntok.clamp(min=1)
            if self.training:
                MTPLossLoggingHelper.save_loss_to_tracker(loss, i, n)
            main_hidden = MTPLossAutoScaler.apply(main_hidden, self.mtp_loss_scaling_factor * loss / n)
        return main_hidden


def get_mtp_layer_spec(transformer_layer_spec: ModuleSpec, norm=None, column_linear=None) -> ModuleSpec:
    from ..parallel.tp_layers import ColumnParallelLinear
    from .norms import FusedLayerNorm

    norm = norm or FusedLayerNorm
    return ModuleSpec(module=MultiTokenPredictionLayer,
                      submodules=MultiTokenPredictionLayerSubmodules(
                          enorm=norm, hnorm=norm, eh_proj=column_linear or ColumnParallelLinear,
                          transformer_layer=transformer_layer_spec, layer_norm=norm))


def get_mtp_block_spec(config, transformer_layer_spec: ModuleSpec) -> Optional[ModuleSpec]:
    n = config.mtp_num_layers or 0
    if n <= 0:
        return None
    ls = get_mtp_layer_spec(transformer_layer_spec)
    return ModuleSpec(module=MultiTokenPredictionBlock,
                      submodules=MultiTokenPredictionBlockSubmodules(layer_specs=[ls] * n))


def get_mtp_num_layers_to_build(config) -> int:
    """MTP layers live on the last pipeline stage only."""
    from ..parallel import state as ps

    return int(config.mtp_num_layers or 0) if ps.is_pipeline_last_stage(ignore_virtual=True) else 0


def get_mtp_layer_offset(config) -> int:
    """All MTP layers are on one stage: their local index is their global index."""
    return 0


def tie_word_embeddings_state_dict(sharded_state_dict: dict, word_emb_weight, word_emb_weight_key: str) -> None:
    """The MTP block's embedding is the model's: store it under the model's key (replica of the first stage's copy)."""
    from ..ckpt.sharded import make_tp_sharded_tensor_for_checkpoint
    from ..parallel import state as ps

    mtp_key = f"{word_emb_weight_key.rsplit('embedding.', 1)[0]}mtp_word_emb.weight"
    sharded_state_dict.pop(mtp_key, None)
    rid = (1, 0, ps.get_data_parallel_rank(with_context_parallel=True))      # not the main replica: stage 0 writes the tensor
    sharded_state_dict[mtp_key] = make_tp_sharded_tensor_for_checkpoint(word_emb_weight, word_emb_weight_key, 0, replica_id=rid)


def tie_output_layer_state_dict(sharded_state_dict: dict, output_layer_weight, output_layer_weight_key: str) -> None:
    from ..ckpt.sharded import make_tp_sharded_tensor_for_checkpoint
    from ..parallel import state as ps

    mtp_key = f"{output_layer_weight_key.rsplit('output_layer.', 1)[0]}mtp_output_layer.weight"
    sharded_state_dict.pop(mtp_key, None)
    rid = (1, 0, ps.get_data_parallel_rank(with_context_parallel=True))
    sharded_state_dict[mtp_key] = make_tp_sharded_tensor_for_checkpoint(output_layer_weight, output_layer_weight_key, 0, replica_id=rid)
