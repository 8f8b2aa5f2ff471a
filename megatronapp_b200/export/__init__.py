from .data_type import DataType  # noqa: F401
from .export_config import ExportConfig  # noqa: F401
from .model_type import ModelType  # noqa: F401
from .trtllm import DistributedTRTLLMConverter, TRTLLMHelper, export_trtllm_checkpoint  # noqa: F401
from .trtllm_engine_builder import TRTLLMEngineBuilder  # noqa: F401
