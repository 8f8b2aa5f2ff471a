"""Reference import path ``core/export/trtllm/trtllm_helper.py``."""
from . import TRTLLMHelper, export_trtllm_checkpoint  # noqa: F401
