from .distillation import LogitsKLLoss, HiddenStateCosineLoss, DistillationLossBalancer, distillation_forward  # noqa: F401
