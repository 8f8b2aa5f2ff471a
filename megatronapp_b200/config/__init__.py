from .transformer_config import (  # noqa: F401
    MLATransformerConfig, ModelParallelConfig, TransformerConfig, init_method_normal,
    scaled_init_method_normal,
)
