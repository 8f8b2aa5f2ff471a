"""Vision -> language projection (reference ``models/vision/multimodal_projector.py``)."""
from __future__ import annotations

from ...transformer.mlp import MLP, MLPSubmodules
from ...transformer.module import MegatronModule
from ...transformer.spec_utils import build_module


class MultimodalProjector(MegatronModule):
    def __init__(self, config, submodules: MLPSubmodules, projector_type: str, input_size: int):
        super().__init__(config)
        self.projector_type = projector_type
        if projector_type == "mlp":
            self.encoder = MLP(config=config, submodules=submodules, input_size=input_size)
        elif projector_type == "affine":
            self.encoder = build_module(submodules.linear_fc1, input_size, config.hidden_size, config=config,
                                        init_method=config.init_method, gather_output=True, bias=config.add_bias_linear,
                                        skip_bias_add=True, is_expert=False)
        else:
            raise ValueError(f"unsupported multimodal projector type {projector_type!r}")

    def forward(self, hidden_states):
        out, bias = self.encoder(hidden_states)
        return out + bias if bias is not None else out
