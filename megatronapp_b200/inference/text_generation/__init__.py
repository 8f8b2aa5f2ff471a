"""The reference's function-style generation API (``megatron/inference/text_generation/api.py``): ``generate`` /
``generate_and_post_process`` / ``beam_search`` / ``beam_search_and_post_process`` over a bare model, as used by its REST
server, its evaluation tasks and user scripts that predate the engine classes."""
from .api import beam_search, beam_search_and_post_process, generate, generate_and_post_process  # noqa: F401
