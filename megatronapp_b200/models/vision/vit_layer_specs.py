"""Layer specs of the vision transformers (reference ``core/models/vision/vit_layer_specs.py``): bidirectional self
attention (``no_mask``), LayerNorm, GELU MLP.  The TE and local variants of the reference collapse into one here -- the
linears are this framework's tcgen05 tensor-parallel layers either way."""
from ...transformer.mlp import MLP, MLPSubmodules
from ...transformer.spec_utils import ModuleSpec
from ...parallel.tp_layers import ColumnParallelLinear, RowParallelLinear
from .clip_vit_model import get_vit_layer_spec


def get_vit_layer_with_transformer_engine_spec() -> ModuleSpec:
    return get_vit_layer_spec()


def get_vit_layer_with_local_spec() -> ModuleSpec:
    return get_vit_layer_spec()


def get_vit_mlp_module_spec(use_te: bool = True) -> ModuleSpec:
    return ModuleSpec(module=MLP, submodules=MLPSubmodules(linear_fc1=ColumnParallelLinear, linear_fc2=RowParallelLinear))
