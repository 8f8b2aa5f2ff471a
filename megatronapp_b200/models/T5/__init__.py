from .t5_model import T5Model, get_t5_decoder_block_spec, get_t5_encoder_block_spec  # noqa: F401
