from .distillation import (DistillationLossBalancer, DistillationState, HiddenStateCosineLoss, LogitsAndIntermediatesLossBalancer,  # noqa: F401
                           LogitsKLLoss, MSELoss, ProjectionLayer, attach_teacher, distillation_forward,
                           load_distillation_config, teacher_args_from_checkpoint)
