"""Zero-overhead MegaScope hook points called from the model code.

When no capture registry / disturbance is installed every hook is a single attribute check.
See ``scope/capture.py`` (flags, compressors, TP re-assembly) and ``scope/disturbance.py``
(noise injection) for the machinery.  Hook sites mirror the reference's:
``attention.py:979-981`` (QKV), ``dot_product_attention.py:168-170`` (probabilities),
``mlp.py:116-118,169-170`` (MLP1/MLP2), ``layers.py:944-963`` (weight perturbation),
``transformer_layer.py:563-565`` (calculation) and ``transformer_block.py:542-544`` (system).
"""
from __future__ import annotations

_CAPTURE = None      # scope.capture.TensorTracers
_DISTURB = None      # scope.disturbance.Disturbance


def install(capture=None, disturbance=None):
    global _CAPTURE, _DISTURB
    _CAPTURE, _DISTURB = capture, disturbance


def capture_active() -> bool:
    return _CAPTURE is not None and _CAPTURE.any_active()


def wants(flag, layer_number) -> bool:
    return _CAPTURE is not None and _CAPTURE.wants(flag, layer_number)


def tik_tensor(flag, layer_number, tensor):
    if _CAPTURE is not None and _CAPTURE.wants(flag, layer_number):
        _CAPTURE.tik_tensor((layer_number, flag), tensor)


def tik_result(*args, **kw):
    if _CAPTURE is not None:
        _CAPTURE.tik_result(*args, **kw)


def tik_end():
    if _CAPTURE is not None:
        _CAPTURE.tik_end()


def perturb_weight(weight, module=None):
    if _DISTURB is None or not _DISTURB.weight_perturbation:
        return weight
    return _DISTURB.apply("weight", weight)


def perturb_calculation(x):
    if _DISTURB is None or not _DISTURB.calculation_perturbation:
        return x
    return _DISTURB.apply("calculation", x)


def perturb_system(x):
    if _DISTURB is None or not _DISTURB.system_perturbation:
        return x
    return _DISTURB.apply("system", x)
