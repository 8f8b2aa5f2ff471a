from .spec_utils import ModuleSpec, build_module  # noqa: F401
from .module import MegatronModule, Float16Module  # noqa: F401
from .transformer_layer import TransformerLayer, TransformerLayerSubmodules  # noqa: F401
from .transformer_block import TransformerBlock, TransformerBlockSubmodules  # noqa: F401
from .attention import SelfAttention, SelfAttentionSubmodules, CrossAttention, CrossAttentionSubmodules, DotProductAttention  # noqa: F401
from .mlp import MLP, MLPSubmodules  # noqa: F401
from ..config.transformer_config import TransformerConfig, MLATransformerConfig  # noqa: F401
