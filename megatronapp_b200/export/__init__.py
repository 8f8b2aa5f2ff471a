from .trtllm import TRTLLMHelper, export_trtllm_checkpoint  # noqa: F401
