"""Process-wide singletons of the training harness (args, tokenizer, timers, writers, tracer).

Parity: reference ``training/global_vars.py`` (``get_args`` :27, ``get_timers`` :61,
``get_tracer`` :71, ``set_global_variables`` :88).  Only the *harness* reads these -- the core
(parallel/, transformer/, models/) never calls ``get_args()``.
"""
from __future__ import annotations

import os
from typing import Optional

from ..scan.tracer import get_tracer as _get_tracer
from ..utils.timers import Timers

_ARGS = None
_TOKENIZER = None
_TB_WRITER = None
_WANDB_WRITER = None
_TIMERS: Optional[Timers] = None
_SIGNAL_HANDLER = None


def get_args():
    if _ARGS is None:
        raise RuntimeError("args is not initialized")
    return _ARGS


def set_args(args):
    global _ARGS
    _ARGS = args


def get_tokenizer():
    return _TOKENIZER


class _ScalarFanOut:
    """What ``training_log`` writes scalars to: TensorBoard and / or Weights & Biases behind one ``add_scalar``."""

    def __init__(self, tb, wandb_run):
        self.tb, self.wandb = tb, wandb_run

    def add_scalar(self, name, value, step):
        if self.tb is not None:
            self.tb.add_scalar(name, value, step)
        if self.wandb is not None:
            self.wandb.log({name: value}, step=int(step))

    def add_text(self, *a, **k):
        if self.tb is not None:
            self.tb.add_text(*a, **k)

    def flush(self):
        if self.tb is not None:
            self.tb.flush()


def get_tensorboard_writer():
    if _TB_WRITER is None and _WANDB_WRITER is None:
        return None
    return _ScalarFanOut(_TB_WRITER, _WANDB_WRITER)


def get_wandb_writer():
    return _WANDB_WRITER


def get_timers() -> Timers:
    if _TIMERS is None:
        raise RuntimeError("timers is not initialized")
    return _TIMERS


def get_tracer():
    return _get_tracer()


def get_signal_handler():
    return _SIGNAL_HANDLER


def set_global_variables(args, build_tokenizer: bool = True):
    global _ARGS, _TOKENIZER, _TIMERS, _TB_WRITER, _SIGNAL_HANDLER, _WANDB_WRITER
    _ARGS = args
    if getattr(args, "logging_level", None) is not None:
        import logging

        logging.getLogger().setLevel(int(args.logging_level))
    from ..utils.microbatches import init_num_microbatches_calculator

    init_num_microbatches_calculator(args.rank, args.rampup_batch_size, args.global_batch_size,
                                     args.micro_batch_size, args.effective_data_parallel_size,
                                     args.decrease_batch_size_if_needed)
    if build_tokenizer:
        from ..data.tokenizer import build_tokenizer as _bt

        _TOKENIZER = _bt(args)
    _TIMERS = Timers(args.timing_log_level, args.timing_log_option)
    if args.tensorboard_dir and args.rank == args.world_size - 1:
        try:
            from torch.utils.tensorboard import SummaryWriter

            _TB_WRITER = SummaryWriter(log_dir=args.tensorboard_dir, max_queue=args.tensorboard_queue_size)
        except Exception:
            print("WARNING: TensorBoard writing requested but is not available", flush=True)
    if getattr(args, "wandb_project", None) and args.rank == args.world_size - 1:
        # reference global_vars._set_wandb_writer: last rank, run directory under --wandb-save-dir (or <save>/wandb);
        # WANDB_MODE=offline keeps it local (no network on the training nodes)
        if not getattr(args, "wandb_exp_name", None):
            raise ValueError("please specify the wandb experiment name (--wandb-exp-name)")
        try:
            import os

            import wandb

            save_dir = args.wandb_save_dir or os.path.join(args.save or ".", "wandb")
            os.makedirs(save_dir, exist_ok=True)
            wandb.init(dir=save_dir, name=args.wandb_exp_name, project=args.wandb_project,
                       config={k: (v if isinstance(v, (int, float, str, bool, type(None))) else str(v)) for k, v in vars(args).items()})
            _WANDB_WRITER = wandb
        except ImportError:
            print("WARNING: --wandb-project given but the wandb package is not available", flush=True)
    tracer = _get_tracer()
    tracer.configure(args.trace, args.trace_dir, args.trace_interval, args.continuous_trace_iterations,
                     args.trace_granularity)
    if getattr(args, "adlr_autoresume", False):
        from .autoresume import build_autoresume

        build_autoresume()
    if args.exit_signal_handler:
        from .signals import DistributedSignalHandler

        _SIGNAL_HANDLER = DistributedSignalHandler().__enter__()


def unset_global_variables():
    global _ARGS, _TOKENIZER, _TIMERS, _TB_WRITER, _SIGNAL_HANDLER, _WANDB_WRITER
    if _WANDB_WRITER is not None:
        _WANDB_WRITER.finish()
    _ARGS = _TOKENIZER = _TIMERS = _TB_WRITER = _SIGNAL_HANDLER = _WANDB_WRITER = None
    from ..utils.microbatches import destroy_num_microbatches_calculator

    destroy_num_microbatches_calculator()


# -- further accessors of the reference's ``global_vars.py`` --------------------------------------------------------------------
def get_one_logger():
    """The reference's optional NVIDIA-internal telemetry client: never configured here."""
    return None


def get_adlr_autoresume():
    from .autoresume import get_adlr_autoresume as impl

    return impl()


def rebuild_tokenizer(args):
    global _TOKENIZER
    from ..data.tokenizer import build_tokenizer as _bt

    _TOKENIZER = _bt(args)
    return _TOKENIZER


def destroy_global_vars():
    unset_global_variables()
