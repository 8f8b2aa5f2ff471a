"""Model-parallel RNG streams and activation recomputation.

Parity: reference ``megatron/core/tensor_parallel/random.py`` (``CudaRNGStatesTracker`` :124,
``model_parallel_cuda_manual_seed`` :319 with tp seed = seed + 2718 + tp_rank,
``CheckpointFunction`` :391, ``CheckpointWithoutOutput`` :495).  Works on CPU too (the
"device" generator is then the default CPU generator) so gloo runs are reproducible.
"""
from __future__ import annotations

import contextlib
from typing import Dict

import torch

from . import state as ps

_MODEL_PARALLEL_RNG_TRACKER_NAME = "model-parallel-rng"
_EXPERT_PARALLEL_RNG_TRACKER_NAME = "expert-parallel-rng"
_DATA_PARALLEL_RNG_TRACKER_NAME = "data-parallel-rng"


def _get_state():
    return torch.cuda.get_rng_state() if torch.cuda.is_available() else torch.get_rng_state()


def _set_state(s):
    if torch.cuda.is_available():
        torch.cuda.set_rng_state(s)
    else:
        torch.set_rng_state(s)


def _seed(seed: int):
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed)
    else:
        torch.manual_seed(seed)


class CudaRNGStatesTracker:
    def __init__(self):
        self.states_: Dict[str, torch.Tensor] = {}
        self.seeds_ = set()

    def reset(self):
        self.states_, self.seeds_ = {}, set()

    def get_states(self):
        return dict(self.states_)

    def set_states(self, states):
        self.states_ = states

    def add(self, name: str, seed: int):
        if seed in self.seeds_:
            raise Exception(f"seed {seed} already exists")
        if name in self.states_:
            raise Exception(f"cuda rng state {name} already exists")
        self.seeds_.add(seed)
        orig = _get_state()
        _seed(seed)
        self.states_[name] = _get_state()
        _set_state(orig)

    @contextlib.contextmanager
    def fork(self, name: str = _MODEL_PARALLEL_RNG_TRACKER_NAME):
        if name not in self.states_:
            raise Exception(f"cuda rng state {name} is not added")
        orig = _get_state()
        _set_state(self.states_[name])
        try:
            yield
        finally:
            self.states_[name] = _get_state()
            _set_state(orig)


_TRACKER = CudaRNGStatesTracker()


def get_cuda_rng_tracker() -> CudaRNGStatesTracker:
    return _TRACKER


def get_expert_parallel_rng_tracker_name():
    return _EXPERT_PARALLEL_RNG_TRACKER_NAME


def get_data_parallel_rng_tracker_name():
    return _DATA_PARALLEL_RNG_TRACKER_NAME


def model_parallel_cuda_manual_seed(seed: int) -> None:
    """default stream: ``seed`` (same across TP, differs per dp in dropout outside TP regions
    only through data); tensor-parallel stream: ``seed + 2718 + tp_rank``; expert stream:
    ``seed + 1024 + 100*ep_rank + tp_rank``."""
    offset = seed + 2718
    tp_seed = offset + ps.get_tensor_model_parallel_rank()
    _TRACKER.reset()
    _seed(seed)
    torch.manual_seed(seed)
    _TRACKER.add(_DATA_PARALLEL_RNG_TRACKER_NAME, seed + 7919)
    _TRACKER.add(_MODEL_PARALLEL_RNG_TRACKER_NAME, tp_seed)
    ep_seed = seed + 1024 + 100 * ps.get_expert_model_parallel_rank() + ps.get_expert_tensor_parallel_rank()
    if ep_seed in _TRACKER.seeds_:
        ep_seed += 104729
    _TRACKER.add(_EXPERT_PARALLEL_RNG_TRACKER_NAME, ep_seed)


class CheckpointFunction(torch.autograd.Function):
    """Recompute ``run_function`` in backward with the forward RNG states restored."""

    @staticmethod
    def forward(ctx, run_function, distribute_saved_activations, *args):
        ctx.run_function = run_function
        ctx.fwd_cpu_rng = torch.get_rng_state()
        ctx.fwd_dev_rng = _get_state()
        ctx.fwd_tracker = get_cuda_rng_tracker().get_states()
        with torch.no_grad():
            outputs = run_function(*args)
        ctx.distribute = bool(distribute_saved_activations) and ps.get_tensor_model_parallel_world_size() > 1
        if ctx.distribute:
            # shard the first input across the TP group (reference random.py:417-432)
            a0 = args[0]
            ctx.input_0_shape = a0.shape
            flat = a0.detach().reshape(-1)
            ws, rk = ps.get_tensor_model_parallel_world_size(), ps.get_tensor_model_parallel_rank()
            n = flat.numel() // ws
            args = (flat[rk * n:(rk + 1) * n].clone(),) + tuple(args[1:])
        ctx.tensor_idx = [i for i, a in enumerate(args) if torch.is_tensor(a)]
        ctx.others = [None if torch.is_tensor(a) else a for a in args]
        ctx.save_for_backward(*[a for a in args if torch.is_tensor(a)])
        return outputs

    @staticmethod
    def backward(ctx, *grads):
        if not torch.autograd._is_checkpoint_valid():
            raise RuntimeError("checkpointing is not compatible with .grad(); use .backward()")
        saved = list(ctx.saved_tensors)
        inputs = list(ctx.others)
        for i, t in zip(ctx.tensor_idx, saved):
            inputs[i] = t
        if ctx.distribute:
            import torch.distributed as dist

            shard = inputs[0]
            ws = ps.get_tensor_model_parallel_world_size()
            full = torch.empty(shard.numel() * ws, dtype=shard.dtype, device=shard.device)
            dist.all_gather_into_tensor(full, shard.contiguous(), group=ps.get_tensor_model_parallel_group())
            inputs[0] = full.view(ctx.input_0_shape)
        bwd_cpu, bwd_dev, bwd_tr = torch.get_rng_state(), _get_state(), get_cuda_rng_tracker().get_states()
        torch.set_rng_state(ctx.fwd_cpu_rng)
        _set_state(ctx.fwd_dev_rng)
        get_cuda_rng_tracker().set_states(ctx.fwd_tracker)
        detached = [a.detach().requires_grad_(a.requires_grad) if torch.is_tensor(a) else a for a in inputs]
        with torch.enable_grad():
            outputs = ctx.run_function(*detached)
        torch.set_rng_state(bwd_cpu)
        _set_state(bwd_dev)
        get_cuda_rng_tracker().set_states(bwd_tr)
        if torch.is_tensor(outputs):
            outputs = (outputs,)
        pairs = [(o, g) for o, g in zip(outputs, grads) if torch.is_tensor(o) and o.requires_grad and g is not None]
        if pairs:
            torch.autograd.backward([o for o, _ in pairs], [g for _, g in pairs])
        in_grads = tuple(a.grad if torch.is_tensor(a) else None for a in detached)
        return (None, None) + in_grads


def checkpoint(function, distribute_saved_activations, *args):
    return CheckpointFunction.apply(function, distribute_saved_activations, *args)


class CheckpointWithoutOutput:
    """Recompute a cheap module (norm / activation) instead of saving its output: the output
    storage is released after its consumer ran and re-materialised right before backward."""

    def __init__(self):
        self.run_function = None
        self.inputs = None
        self.outputs = None
        self.rng = None

    def checkpoint(self, run_function, *args):
        self.run_function = run_function
        self.rng = (torch.get_rng_state(), _get_state(), get_cuda_rng_tracker().get_states())
        self.inputs = args
        with torch.no_grad():
            out = run_function(*[a.detach() if torch.is_tensor(a) else a for a in args])
        # make the output a leaf that we will re-fill in backward
        self.outputs = out
        self._proxy = out.detach().requires_grad_(True)
        return self._proxy

    def discard_output_and_register_recompute(self, hook_tensor: torch.Tensor):
        self._proxy.untyped_storage().resize_(0) if self._proxy.device.type == "cuda" else None
        if hook_tensor.requires_grad:
            hook_tensor.register_hook(self._recompute)

    def _recompute(self, _grad):
        cpu, dev, tr = torch.get_rng_state(), _get_state(), get_cuda_rng_tracker().get_states()
        torch.set_rng_state(self.rng[0]); _set_state(self.rng[1]); get_cuda_rng_tracker().set_states(self.rng[2])
        with torch.enable_grad():
            det = [a.detach().requires_grad_(a.requires_grad) if torch.is_tensor(a) else a for a in self.inputs]
            out = self.run_function(*det)
        torch.set_rng_state(cpu); _set_state(dev); get_cuda_rng_tracker().set_states(tr)
        if self._proxy.device.type == "cuda":
            self._proxy.untyped_storage().resize_(out.untyped_storage().size())
            self._proxy.data.copy_(out.detach())
        self._recomputed = (det, out)
        return _grad

    def backward_through(self):
        """Propagate ``proxy.grad`` through the recomputed graph into the original inputs."""
        det, out = self._recomputed
        torch.autograd.backward(out, self._proxy.grad)
        for orig, d in zip(self.inputs, det):
            if torch.is_tensor(orig) and orig.requires_grad and d.grad is not None:
                orig.backward(d.grad)


def initialize_rng_tracker(use_te_rng_tracker: bool = False, inference_rng_tracker: bool = False, use_cudagraphable_rng: bool = False,
                           force_reset: bool = False):
    """Create (or with ``force_reset`` re-create) the global RNG-state tracker.  There is one tracker implementation here: its
    states are plain generator states, which is what CUDA-graph capture and inference need as well."""
    global _TRACKER
    if force_reset or globals().get("_TRACKER") is None:
        _TRACKER = CudaRNGStatesTracker()
    return _TRACKER


def get_all_rng_states() -> dict:
    """``{name: state}`` of the tracker (empty before ``model_parallel_cuda_manual_seed``)."""
    return get_cuda_rng_tracker().get_states()


CheckpointWithoutOutputFunction = CheckpointFunction
