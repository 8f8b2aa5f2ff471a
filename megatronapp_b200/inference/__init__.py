from .core import (DynamicInferenceEngine, GPTInferenceWrapper, InferenceParams, InferenceRequest, SamplingParams,  # noqa: F401
                   Scheduler, StaticInferenceContext, StaticInferenceEngine, Status, TextGenerationController,
                   sample_from_logits)
