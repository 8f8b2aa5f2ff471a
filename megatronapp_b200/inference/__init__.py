from .core import (GPTInferenceWrapper, InferenceParams, InferenceRequest, SamplingParams,  # noqa: F401
                   Scheduler, StaticInferenceContext, StaticInferenceEngine, Status, T5InferenceWrapper, TextGenerationController,
                   VLMInferenceWrapper, sample_from_logits)
from .dynamic import BlockAllocator, DynamicInferenceContext, DynamicInferenceEngine  # noqa: F401
from .async_stream import AsyncStream, TokenStream  # noqa: F401,E402
