from .retro_model import RetroConfig, RetroModel, get_retro_decoder_block_spec  # noqa: F401
