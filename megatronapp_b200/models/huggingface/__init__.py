"""HuggingFace towers as Megatron modules (reference ``core/models/huggingface``): a vision encoder or a language model that
exists only as a ``transformers`` checkpoint can be plugged into LLaVA (``--vision-model-type hf://<path>`` /
``--language-model-type hf://<path>``) without converting it first."""
from .module import AutoHuggingFaceModel, HuggingFaceModule, build_hf_model, get_hf_model_type  # noqa: F401
