"""Process initialisation: args -> globals -> torch.distributed -> parallel state -> seeds.

Parity: reference ``training/initialize.py`` (``initialize_megatron`` :37-166,
``_initialize_distributed`` :295-377 incl. the warm-up all-reduce, ``_set_random_seed`` :399-421).
No runtime ``make`` of helpers (built at install time), no hard import of RDMA extensions
(SURVEY Q1), no ``rm -rf /dev/shm/*`` (Q2), and the gloo/CPU mode is first-class.
"""
from __future__ import annotations

import datetime
import os
import random

import numpy as np
import torch
import torch.distributed as dist

from ..parallel import p2p
from ..parallel import state as ps
from ..parallel.random import model_parallel_cuda_manual_seed
from .arguments import parse_args, validate_args
from .global_vars import set_global_variables


def _apply_deterministic_mode(args):
    """``--deterministic-mode``: bitwise-reproducible runs (reference arguments.py / initialize.py): deterministic
    torch algorithms, no TF32 surprises, NCCL algorithm pinned.  Our own kernels are deterministic by
    construction (no atomics in GEMM / attention / norm reductions; the fused reduce-scatter reduces in rank order)."""
    if not getattr(args, "deterministic_mode", False):
        return
    import os

    import torch

    os.environ.setdefault("NCCL_ALGO", "Ring")
    os.environ.setdefault("CUBLAS_WORKSPACE_CONFIG", ":4096:8")
    torch.use_deterministic_algorithms(True, warn_only=True)
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False


def initialize_megatron(extra_args_provider=None, args_defaults=None, ignore_unknown_args=False,
                        allow_no_cuda=True, skip_mpu_initialization=False, args_list=None, parsed_args=None):
    args = parsed_args if parsed_args is not None else parse_args(extra_args_provider, ignore_unknown_args, args_list)
    if args.use_checkpoint_args and args.load:
        from .checkpointing import load_args_from_checkpoint

        load_args_from_checkpoint(args)
    if not torch.cuda.is_available():
        if not allow_no_cuda:
            raise RuntimeError("this entry point requires CUDA")
        args.distributed_backend = "gloo"
        if "OMP_NUM_THREADS" not in os.environ:
            torch.set_num_threads(max(1, min(4, (os.cpu_count() or 2) // max(1, args.world_size))))
    validate_args(args, args_defaults or {})
    set_global_variables(args)
    _apply_deterministic_mode(args)
    if not skip_mpu_initialization:
        _initialize_distributed(args)
        _set_random_seed(args.seed, args.data_parallel_random_init)
    if args.rank == 0:
        _print_args(args)
    return args


def _initialize_distributed(args):
    n_dev = torch.cuda.device_count() if torch.cuda.is_available() else 0
    if not dist.is_initialized():
        if n_dev > 0:
            torch.cuda.set_device(args.local_rank % n_dev)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        kw = {}
        if args.distributed_backend == "nccl" and n_dev > 0:
            kw["device_id"] = torch.device("cuda", torch.cuda.current_device())
        if args.world_size > 1 or "RANK" in os.environ:
            dist.init_process_group(backend=args.distributed_backend, world_size=args.world_size, rank=args.rank,
                                    timeout=datetime.timedelta(minutes=args.distributed_timeout_minutes), **kw)
            # warm-up collective so communicator setup is not billed to the first iteration
            t = torch.zeros(1, device="cuda" if (args.distributed_backend == "nccl" and n_dev > 0) else "cpu")
            dist.all_reduce(t)
    if ps.model_parallel_is_initialized():
        return
    split = args.pipeline_model_parallel_split_rank or args.encoder_pipeline_model_parallel_size or None
    if split is None and getattr(args, "model_type_hint", None) == "encoder_and_decoder" \
            and args.pipeline_model_parallel_size > 1:
        split = args.pipeline_model_parallel_size // 2          # half of the stages each, like the reference's T5 recipes
    args.pipeline_model_parallel_split_rank = split
    ps.initialize_model_parallel(
        args.tensor_model_parallel_size, args.pipeline_model_parallel_size,
        args.virtual_pipeline_model_parallel_size, split,
        context_parallel_size=args.context_parallel_size,
        hierarchical_context_parallel_sizes=args.hierarchical_context_parallel_sizes,
        expert_model_parallel_size=args.expert_model_parallel_size,
        num_distributed_optimizer_instances=args.num_distributed_optimizer_instances,
        expert_tensor_parallel_size=args.expert_tensor_parallel_size,
        distributed_timeout_minutes=args.distributed_timeout_minutes,
        nccl_communicator_config_path=args.nccl_communicator_config_path,
        order="tp-cp-ep-dp-pp" if not args.use_tp_pp_dp_mapping else "tp-cp-ep-pp-dp",
        forward_backward_disaggregating=args.forward_backward_disaggregating,
        ignore_forward_tensor_parallel=bool(getattr(args, "ignore_forward_tensor_parallel", False)),
        world_size=args.world_size, rank=args.rank)
    if dist.is_initialized():
        p2p.build_channels()
    if args.rank == 0:
        print(f"> initialized {ps.get_topology().describe()}", flush=True)


def _set_random_seed(seed_, data_parallel_random_init=False):
    if seed_ is None or seed_ <= 0:
        raise ValueError(f"seed ({seed_}) should be a positive integer")
    seed = seed_ + 100 * ps.get_pipeline_model_parallel_rank()
    if data_parallel_random_init:
        seed = seed + 10 * ps.get_data_parallel_rank()
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    model_parallel_cuda_manual_seed(seed)


def _print_args(args):
    print("------------------------ arguments ------------------------", flush=True)
    for k in sorted(vars(args)):
        print(f"  {k} {'.' * max(1, 48 - len(k))} {getattr(args, k)}", flush=True)
    print("-------------------- end of arguments ---------------------", flush=True)


# -- further entry points of the reference's ``initialize.py`` -----------------------------------------------------------------
def write_args_to_tensorboard():
    """Every argument as a text entry of the TensorBoard run (reference ``write_args_to_tensorboard``)."""
    from .global_vars import get_args, get_tensorboard_writer

    args, writer = get_args(), get_tensorboard_writer()
    if writer is not None:
        for name in sorted(vars(args)):
            writer.add_text(name, str(getattr(args, name)), global_step=args.iteration if hasattr(args, "iteration") else 0)


def set_jit_fusion_options():
    """The reference configures the TorchScript / nvFuser fusers for its jit-compiled bias-GELU / bias-dropout-add here.  Those
    ops are hand-written kernels in this framework and nothing is traced, so there is nothing to configure."""
    return None


def setup_logging() -> None:
    """Logging level of the root logger: ``--logging-level``, else the ``MEGATRON_LOGGING_LEVEL`` environment variable."""
    import logging
    import os

    from .global_vars import get_args

    level = getattr(get_args(), "logging_level", None)
    if level is None and os.environ.get("MEGATRON_LOGGING_LEVEL"):
        level = int(os.environ["MEGATRON_LOGGING_LEVEL"])
    if level is not None:
        logging.getLogger().setLevel(int(level))
