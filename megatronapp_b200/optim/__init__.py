from .optimizer import (  # noqa: F401
    ChainedOptimizer, ConstantGradScaler, DistributedOptimizer, DynamicGradScaler, FP32Optimizer,
    Float16OptimizerWithFloat16Params, MegatronOptimizer, OptimizerConfig, get_megatron_optimizer,
)
from .scheduler import OptimizerParamScheduler  # noqa: F401
