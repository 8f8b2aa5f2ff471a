"""Base class of the vision towers (reference ``core/models/common/vision_module/vision_module.py``)."""
from ....transformer.module import MegatronModule


class VisionModule(MegatronModule):
    """Marker base: image encoders take pixel tensors, are never tied to an output layer and are usually frozen or trained
    with a learning-rate multiplier of their own."""

    share_embeddings_and_output_weights = False

    def __init__(self, config):
        super().__init__(config)
