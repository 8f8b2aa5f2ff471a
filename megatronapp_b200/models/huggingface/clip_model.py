"""HF vision towers (reference ``core/models/huggingface/clip_model.py``): images in, patch embeddings ``[b, patches(+cls), h]``
out -- the interface of ``CLIPViTModel``."""
from .module import HuggingFaceModule, load_auto_model


class _VisionTower(HuggingFaceModule):
    def __init__(self, config):
        super().__init__(config)
        m = load_auto_model(config.huggingface_model_name_or_path)
        self.model = getattr(m, "vision_model", m)          # CLIPModel / SiglipModel carry a text tower we do not need
        hc = getattr(self.model, "config", None)
        self.hidden_size = getattr(hc, "hidden_size", None)

    def forward(self, images, **kw):
        out = self.model(pixel_values=images, **kw)
        return out.last_hidden_state if hasattr(out, "last_hidden_state") else out[0]


class ClipHuggingFaceModel(_VisionTower):
    add_class_token, class_token_len = True, 1


class SiglipHuggingFaceModel(_VisionTower):
    add_class_token, class_token_len = False, 0
