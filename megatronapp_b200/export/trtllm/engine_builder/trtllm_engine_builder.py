"""Reference import path ``core/export/trtllm/engine_builder/trtllm_engine_builder.py``."""
from ...trtllm_engine_builder import EnginePlan, TRTLLMEngineBuilder, resolve_num_tokens  # noqa: F401
