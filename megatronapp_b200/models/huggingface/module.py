"""Wrapper base + factory (reference ``core/models/huggingface/module.py``).  ``model_path`` is ``hf://<name or directory>``;
a directory that only holds ``config.json`` builds the architecture with random weights (``from_config``) -- what the
offline tests use -- otherwise ``from_pretrained`` loads the weights."""
from __future__ import annotations

import os

from ...transformer.module import MegatronModule


def _strip(path: str) -> str:
    return path.split("hf://", 1)[1] if path.startswith("hf://") else path


def load_auto_model(path: str, causal_lm: bool = False):
    from transformers import AutoConfig, AutoModel, AutoModelForCausalLM

    path = _strip(path)
    cls = AutoModelForCausalLM if causal_lm else AutoModel
    has_weights = (not os.path.isdir(path)) or any(f.endswith((".safetensors", ".bin")) for f in os.listdir(path))
    if has_weights:
        return cls.from_pretrained(path)
    return cls.from_config(AutoConfig.from_pretrained(path))


class HuggingFaceModule(MegatronModule):
    """A HF model behind the interface the pipeline schedule and the LLaVA container expect."""

    def __init__(self, config):
        super().__init__(config)
        self.input_tensor = None

    def set_input_tensor(self, input_tensor):
        self.input_tensor = input_tensor          # HF towers always run whole on one stage

    def sharded_state_dict(self, prefix="", sharded_offsets=(), metadata=None):
        from ...ckpt.sharded import make_sharded_tensors_for_checkpoint

        # replicated over TP: every tensor is stored once
        return make_sharded_tensors_for_checkpoint(self.state_dict(prefix="", keep_vars=True), prefix, {}, sharded_offsets)


class AutoHuggingFaceModel(HuggingFaceModule):
    def __init__(self, config):
        super().__init__(config)
        self.model = load_auto_model(config.huggingface_model_name_or_path)

    def forward(self, *args, **kwargs):
        return self.model(*args, **kwargs)


def get_hf_model_type(model_path: str) -> str:
    from transformers import AutoConfig

    cfg = AutoConfig.from_pretrained(_strip(model_path))
    arch = ((cfg.architectures or [type(cfg).__name__])[0]).lower()
    for key in ("qwen", "siglip", "clip"):
        if key in arch or key in getattr(cfg, "model_type", ""):
            return key
    raise NotImplementedError(f"unsupported huggingface model {arch}")


def build_hf_model(config, model_path: str) -> HuggingFaceModule:
    kind = get_hf_model_type(model_path)
    config.huggingface_model_name_or_path = model_path
    if kind == "qwen":
        from .qwen_model import QwenHuggingFaceModel

        return QwenHuggingFaceModel(config)
    from .clip_model import ClipHuggingFaceModel, SiglipHuggingFaceModel

    return SiglipHuggingFaceModel(config) if kind == "siglip" else ClipHuggingFaceModel(config)
