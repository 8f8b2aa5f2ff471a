"""Reference import path: the single-device converter is ``TRTLLMHelper.convert`` + ``shard`` + ``split_pipeline``."""
from .. import TRTLLMHelper as SingleDeviceTRTLLMModelWeightsConverter  # noqa: F401
