"""Legacy vision research models (reference ``megatron/legacy/model/vision``): Swin and MiT backbones, DINO
self-distillation, ViT / SegFormer-style inpainting.  Self-contained PyTorch modules; the ViT trunk is the framework's own
``CLIPViTModel`` (tcgen05 GEMMs, flash attention), the hierarchical backbones use windowed / spatially reduced attention."""
from .dino import DINOHead, DINOLoss, DINOPretrainModel  # noqa: F401
from .inpainting import MitInpaintingModel, VitInpaintingModel  # noqa: F401
from .mit_backbone import MixVisionTransformer, mit_b0, mit_b1, mit_b2  # noqa: F401
from .swin_backbone import SwinTransformer, get_swin  # noqa: F401
