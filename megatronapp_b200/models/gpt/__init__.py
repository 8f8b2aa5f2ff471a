from .gpt_model import GPTModel  # noqa: F401
from .gpt_layer_specs import get_gpt_layer_local_spec, get_gpt_layer_with_transformer_engine_spec, get_gpt_decoder_block_spec  # noqa: F401
