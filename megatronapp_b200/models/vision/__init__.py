from .clip_vit_model import CLIPViTModel, get_num_image_embeddings, get_vit_layer_spec  # noqa: F401
from .multimodal_projector import MultimodalProjector  # noqa: F401
