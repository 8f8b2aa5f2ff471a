"""Reference import path: the on-device converter."""
from .. import DistributedTRTLLMConverter as DistributedTRTLLMModelWeightsConverter  # noqa: F401
