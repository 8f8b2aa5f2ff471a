"""Model families the exporter knows (reference ``core/export/model_type.py``) and their TRT-LLM architecture names
(``trt_model_type.py``)."""
from enum import Enum


class ModelType(Enum):
    gpt = "gpt"
    gptnext = "gptnext"
    llama = "llama"
    falcon = "falcon"
    starcoder = "starcoder"
    mixtral = "mixtral"
    gemma = "gemma"
    nemotron_nas = "nemotron_nas"


TRT_MODEL_TYPE_STRING = {
    ModelType.gpt: "GPTForCausalLM", ModelType.gptnext: "GPTForCausalLM", ModelType.starcoder: "GPTForCausalLM",
    ModelType.mixtral: "LlamaForCausalLM", ModelType.llama: "LlamaForCausalLM", ModelType.gemma: "GemmaForCausalLM",
    ModelType.falcon: "FalconForCausalLM", ModelType.nemotron_nas: "DeciLMForCausalLM",
}
