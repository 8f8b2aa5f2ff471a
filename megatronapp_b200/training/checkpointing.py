"""Checkpoint save / load in the Megatron layout.

Parity: reference ``training/checkpointing.py`` -- tracker ``latest_checkpointed_iteration.txt``
(:222-226), directory ``iter_{:07d}`` (:121-124), legacy ``torch`` format
``mp_rank_{tp:02d}[_{pp:03d}][_{ep:03d}]/model_optim_rng.pt`` (+ ``distrib_optim.pt``)
(:152-168) with state-dict keys ``args, checkpoint_version (3.0), iteration, model | model{i},
optimizer, opt_param_scheduler, rng_state, num_floating_point_operations_so_far``
(:729-780); ``torch_dist`` = sharded checkpoint with resharding (``ckpt/sharded.py``);
``--async-save`` hands the CPU copy to a background writer; MegaFBD forces the per-rank format.
"""
from __future__ import annotations

import contextlib
import os
import random
import threading
from typing import List, Optional

import numpy as np
import torch
import torch.distributed as dist

from ..ckpt import sharded, torch_dist
from ..parallel import state as ps
from ..parallel.random import get_cuda_rng_tracker

_CHECKPOINT_VERSION = 3.0
_ASYNC_THREADS: List[threading.Thread] = []
_PENDING: List[dict] = []          # async saves whose tracker update is deferred to maybe_finalize_async_save


def get_checkpoint_tracker_filename(checkpoints_path):
    return os.path.join(checkpoints_path, "latest_checkpointed_iteration.txt")


def get_checkpoint_name(checkpoints_path, iteration, release=False, pipeline_parallel=None, tensor_rank=None,
                        pipeline_rank=None, expert_parallel=None, expert_rank=None, return_base_dir=False,
                        basename="model_optim_rng.pt"):
    directory = "release" if release else "iter_{:07d}".format(iteration)
    base = os.path.join(checkpoints_path, directory)
    if return_base_dir:
        return base
    if pipeline_parallel is None:
        pipeline_parallel = ps.get_pipeline_model_parallel_world_size() > 1
    if tensor_rank is None:
        tensor_rank = ps.get_tensor_model_parallel_rank()
    if pipeline_rank is None:
        pipeline_rank = ps.get_pipeline_model_parallel_rank()
    if expert_parallel is None:
        expert_parallel = ps.get_expert_model_parallel_world_size() > 1
    if expert_rank is None:
        expert_rank = ps.get_expert_model_parallel_rank()
    name = f"mp_rank_{tensor_rank:02d}"
    if pipeline_parallel:
        name += f"_{pipeline_rank:03d}"
    if expert_parallel:
        name += f"_{expert_rank:03d}"
    return os.path.join(base, name, basename)


def read_metadata(tracker_filename):
    with open(tracker_filename) as f:
        s = f.read().strip()
    if s == "release":
        return 0, True
    return int(s), False


def get_rng_state():
    st = {"random_rng_state": random.getstate(), "np_rng_state": np.random.get_state(),
          "torch_rng_state": torch.get_rng_state(),
          "cuda_rng_state": torch.cuda.get_rng_state() if torch.cuda.is_available() else None,
          "rng_tracker_states": get_cuda_rng_tracker().get_states()}
    return [st]


def _unwrap(m):
    while hasattr(m, "module"):
        m = m.module
    return m


def _fsdp_wrapper(m):
    """The ZeRO-3 wrapper around ``m`` (or None): its parameters only exist in full inside ``summon_full_params``."""
    while m is not None:
        if hasattr(m, "summon_full_params"):
            return m
        m = getattr(m, "module", None)
    return None


@contextlib.contextmanager
def _full_params(model, writeback: bool):
    with contextlib.ExitStack() as stack:
        for m in model:
            w = _fsdp_wrapper(m)
            if w is not None:
                stack.enter_context(w.summon_full_params(writeback=writeback))
        yield


def generate_state_dict(args, model, optimizer, opt_param_scheduler, rng_state, iteration=None,
                        num_floating_point_operations_so_far=0, model_sd_fn=None):
    sd = {"args": args, "checkpoint_version": _CHECKPOINT_VERSION}
    if iteration is not None:
        sd["iteration"] = iteration
    fn = model_sd_fn or (lambda m: m.state_dict_for_save_checkpoint())
    if len(model) == 1:
        sd["model"] = fn(model[0])
    else:
        for i, m in enumerate(model):
            sd[f"model{i}"] = fn(m)
    if not getattr(args, "no_save_optim", False):
        if optimizer is not None:
            sd["optimizer"] = optimizer.state_dict()
        if opt_param_scheduler is not None:
            sd["opt_param_scheduler"] = opt_param_scheduler.state_dict()
    if not getattr(args, "no_save_rng", False):
        sd["rng_state"] = rng_state
    sd["num_floating_point_operations_so_far"] = num_floating_point_operations_so_far
    return sd


def _to_cpu(obj):
    if torch.is_tensor(obj):
        return obj.detach().to("cpu", copy=True)
    if isinstance(obj, dict):
        return {k: _to_cpu(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_to_cpu(v) for v in obj)
    return obj


_LOCAL_MGR = None


def local_checkpoint_manager(args):
    """``--non-persistent-ckpt-type local``: the per-rank, node-local checkpoint manager (one per process)."""
    global _LOCAL_MGR
    if _LOCAL_MGR is None or _LOCAL_MGR.dir != args.non_persistent_local_ckpt_dir:
        from .local_checkpoint import LocalCheckpointManager

        _LOCAL_MGR = LocalCheckpointManager(args.non_persistent_local_ckpt_dir, bool(getattr(args, "replication", False)),
                                            getattr(args, "replication_jump", None), getattr(args, "replication_factor", 2))
    return _LOCAL_MGR


def non_persistent_dir(args) -> str:
    """Where ``--non-persistent-ckpt-type global`` checkpoints go: frequent, overwritten-in-place safety checkpoints next
    to (not instead of) the persistent ones (reference checkpointing.py: ``non_persistent_ckpt``)."""
    return args.non_persistent_global_ckpt_dir or os.path.join(args.save or args.load, "non_persistent")


def save_checkpoint(iteration, model, optimizer, opt_param_scheduler, num_floating_point_operations_so_far=0,
                    args=None, checkpointing_context=None, non_persistent_ckpt: bool = False):
    if args is None:
        from .global_vars import get_args

        args = get_args()
    if non_persistent_ckpt and getattr(args, "non_persistent_ckpt_type", None) == "local":
        # every rank keeps its own complete state on node-local storage (no shared file system, no gather)
        chunks = model if isinstance(model, (list, tuple)) else [model]
        if any(_fsdp_wrapper(m) is not None for m in chunks):
            raise NotImplementedError("local non-persistent checkpoints with ZeRO-3 wrappers: use --non-persistent-ckpt-type global")
        sd = generate_state_dict(args, chunks, optimizer, opt_param_scheduler, get_rng_state(), iteration,
                                 num_floating_point_operations_so_far)
        sd["parallel_layout"] = _layout_tag()
        local_checkpoint_manager(args).save(_to_cpu(sd), iteration)
        if not dist.is_initialized() or dist.get_rank() == 0:
            print(f"  successfully saved local checkpoint from iteration {iteration:7d} to {args.non_persistent_local_ckpt_dir}", flush=True)
        return
    if non_persistent_ckpt:
        # same writer, different root; once the new one is complete the previous non-persistent iteration is removed
        import copy
        import shutil

        npd = non_persistent_dir(args)
        old = None
        tr = get_checkpoint_tracker_filename(npd)
        if os.path.isfile(tr):
            old = read_metadata(tr)[0]
        a2 = copy.copy(args)
        a2.save, a2.async_save = npd, False
        save_checkpoint(iteration, model, optimizer, opt_param_scheduler, num_floating_point_operations_so_far, a2)
        if old is not None and old != iteration and (not dist.is_initialized() or dist.get_rank() == 0):
            shutil.rmtree(get_checkpoint_name(npd, old, return_base_dir=True), ignore_errors=True)
        return
    model = model if isinstance(model, (list, tuple)) else [model]
    rank = dist.get_rank() if dist.is_initialized() else 0
    base = get_checkpoint_name(args.save, iteration, return_base_dir=True)
    fmt = args.ckpt_format
    rng_state = get_rng_state()
    pending = []
    if fmt == "torch":
        dp_rank = ps.get_data_parallel_rank(with_context_parallel=True)
        name = get_checkpoint_name(args.save, iteration)
        sharded_opt = optimizer is not None and getattr(optimizer, "sharded", False)
        # one writer per (tp, pp, ep) coordinate: with expert parallelism the experts differ across the dense DP group,
        # so the writer is the rank with EXPERT-data-parallel rank 0 (reference checkpointing.py: same condition)
        writer = (ps.get_expert_data_parallel_rank() == 0) if ps.get_expert_model_parallel_world_size() > 1 else dp_rank == 0
        if writer or sharded_opt or args.forward_backward_disaggregating:
            os.makedirs(os.path.dirname(name), exist_ok=True)
        gathered = None
        if any(_fsdp_wrapper(m) is not None for m in model):
            # ZeRO-3: assembling the full parameters is a collective -- every rank takes part, the writer keeps the result
            gathered = iter([m.state_dict_for_save_checkpoint() for m in model])
        if writer or args.forward_backward_disaggregating:
            sd = generate_state_dict(args, model, None if sharded_opt else optimizer, opt_param_scheduler,
                                     rng_state, iteration, num_floating_point_operations_so_far,
                                     model_sd_fn=(lambda m: next(gathered)) if gathered is not None else None)
            sd = _to_cpu(sd)
            _write(sd, name, args.async_save)
        if sharded_opt and not args.no_save_optim:
            oname = os.path.join(os.path.dirname(name), f"distrib_optim.pt" if dp_rank == 0 else f"distrib_optim_dp{dp_rank:03d}.pt")
            _write(_to_cpu(optimizer.state_dict()), oname, args.async_save)
    else:
        common = {"args": args, "checkpoint_version": _CHECKPOINT_VERSION, "iteration": iteration,
                  "num_floating_point_operations_so_far": num_floating_point_operations_so_far,
                  "opt_param_scheduler": opt_param_scheduler.state_dict() if opt_param_scheduler else None,
                  "rng_state": rng_state, "parallel_layout": _layout_tag()}
        dl = getattr(args, "_dataloader_state_fn", None)
        if dl is not None:
            common["dataloader_state"] = dl()        # reference checkpointing.py:681 (train_data_iterator.save_state)
        with _full_params(model, writeback=False):
            ssd = _model_sharded_state_dict(model, args)
            if fmt == "b200_sharded":
                sharded.save(ssd, base, common)
                if optimizer is not None and not args.no_save_optim:
                    torch.save(_to_cpu(optimizer.state_dict()), os.path.join(base, f"optim_{rank:05d}.pt"))
            else:
                if (optimizer is not None and not args.no_save_optim and hasattr(optimizer, "sharded_common_state")
                        and getattr(optimizer, "supports_model_space_state", lambda: True)()):
                    ssd.update(optimizer.sharded_state_dict(ssd))
                    common["optimizer"] = optimizer.sharded_common_state()
                elif optimizer is not None and not args.no_save_optim:
                    # optimizers without a model-space view (ZeRO-3 wrappers): per-rank state, same-layout restore only
                    os.makedirs(base, exist_ok=True)
                    _write(_to_cpu(optimizer.state_dict()), os.path.join(base, f"optim_{rank:05d}.pt"), args.async_save)
                pending.append(torch_dist.save(ssd, base, common, async_save=bool(args.async_save)))
    _finish_save(args, iteration, pending, defer=bool(args.async_save))


def _layout_tag() -> dict:
    if not ps.model_parallel_is_initialized():
        return {}
    return {"tp": ps.get_tensor_model_parallel_world_size(), "pp": ps.get_pipeline_model_parallel_world_size(),
            "dp": ps.get_data_parallel_world_size(with_context_parallel=True), "ep": ps.get_expert_model_parallel_world_size()}


def _model_sharded_state_dict(model, args) -> dict:
    """Model keys in the reference's layout: homogeneous decoder layers stacked under one key (transformer_block.py:600)
    unless the model is MoE / heterogeneous or ``--ckpt-... non-homogeneous`` was asked for."""
    ssd = {}
    for i, m in enumerate(model):
        prefix = "" if len(model) == 1 else f"model{i}."
        ssd.update(_unwrap(m).sharded_state_dict(prefix=prefix))
    cfg = getattr(_unwrap(model[0]), "config", None)
    homogeneous = (cfg is not None and getattr(cfg, "num_moe_experts", None) is None
                   and not getattr(cfg, "heterogeneous_block_specs", False)
                   and not getattr(args, "hetereogenous_dist_checkpoint", False)
                   and not getattr(args, "ckpt_non_homogeneous_layers", False)
                   and getattr(args, "ckpt_format", "torch_dist") != "b200_sharded")
    if homogeneous:
        ssd = torch_dist.stack_homogeneous_layers(ssd, cfg.num_layers)
    return ssd


def _finish_save(args, iteration, pending, defer: bool):
    """A checkpoint becomes visible only when every rank's data is on disk: write the tracker after the (cross-rank)
    finalize.  With ``--async-save`` this runs from ``maybe_finalize_async_save`` (reference async_utils.py:90-230)."""
    rec = {"args_save": args.save, "iteration": iteration, "pending": pending, "threads": list(_ASYNC_THREADS)}
    if defer:
        _PENDING.append(rec)
        return
    _complete(rec)


def _complete(rec):
    for t in rec["threads"]:
        t.join()
        if t in _ASYNC_THREADS:
            _ASYNC_THREADS.remove(t)
    for p in rec["pending"]:
        p.finalize()
    rank = dist.get_rank() if dist.is_initialized() else 0
    if dist.is_initialized():
        dist.barrier()
    if rank == 0:
        tr = get_checkpoint_tracker_filename(rec["args_save"])
        with open(tr + ".tmp", "w") as f:
            f.write(str(rec["iteration"]))
        os.replace(tr + ".tmp", tr)
        print(f"  successfully saved checkpoint from iteration {rec['iteration']:7d} to {rec['args_save']}", flush=True)
    if dist.is_initialized():
        dist.barrier()


def _atomic_torch_save(obj, path):
    torch.save(obj, path + ".tmp")
    os.replace(path + ".tmp", path)


def _write(obj, path, async_save):
    if async_save:
        t = threading.Thread(target=_atomic_torch_save, args=(obj, path), daemon=False)
        t.start()
        _ASYNC_THREADS.append(t)
    else:
        _atomic_torch_save(obj, path)


def maybe_finalize_async_save(blocking: bool = False):
    """Complete the oldest outstanding async save once EVERY rank's writer has finished (an all-reduce of the local
    "done" flags keeps the ranks in lock step); only then is the tracker file updated."""
    while _PENDING:
        rec = _PENDING[0]
        done = all(not t.is_alive() for t in rec["threads"]) and all(p.ready() for p in rec["pending"])
        if not blocking:
            if dist.is_initialized() and dist.get_world_size() > 1:
                flag = torch.tensor([1 if done else 0], dtype=torch.int32,
                                    device="cuda" if (torch.cuda.is_available() and dist.get_backend() == "nccl") else "cpu")
                dist.all_reduce(flag, op=dist.ReduceOp.MIN)
                done = bool(flag.item())
            if not done:
                return
        _PENDING.pop(0)
        _complete(rec)
        from . import ft_integration

        ft_integration.on_checkpointing_end(is_async_finalization=True)


def load_checkpoint(model, optimizer, opt_param_scheduler, load_arg="load", strict=True, args=None):
    if args is None:
        from .global_vars import get_args

        args = get_args()
    load_dir = getattr(args, load_arg)
    model = model if isinstance(model, (list, tuple)) else [model]
    if getattr(args, "retro_add_retriever", False):
        strict = False        # a GPT checkpoint may initialise a Retro model: the retriever's weights are simply absent (reference :1455)
    if load_dir and getattr(args, "non_persistent_ckpt_type", None) == "global" and not getattr(args, "ckpt_step", None):
        # resume from whichever is newer: the last persistent checkpoint or the non-persistent safety checkpoint
        npd = args.non_persistent_global_ckpt_dir or os.path.join(load_dir, "non_persistent")
        tr_np, tr_p = get_checkpoint_tracker_filename(npd), get_checkpoint_tracker_filename(load_dir)
        if os.path.isfile(tr_np):
            it_np = read_metadata(tr_np)[0]
            it_p = read_metadata(tr_p)[0] if os.path.isfile(tr_p) else -1
            if it_np > it_p:
                load_dir = npd
    tracker = get_checkpoint_tracker_filename(load_dir) if load_dir else None
    local_sd = None
    if (load_arg == "load" and getattr(args, "non_persistent_ckpt_type", None) == "local" and args.non_persistent_local_ckpt_dir
            and not getattr(args, "ckpt_step", None) and not args.finetune):
        # node-local safety checkpoint: used when it is newer than the last persistent one and every rank's state is
        # retrievable (own file or a replica); same parallel layout only
        mgr = local_checkpoint_manager(args)
        it_local = mgr.find_latest()
        it_p = read_metadata(tracker)[0] if (tracker and os.path.isfile(tracker)) else -1
        if it_local > it_p:
            cand = mgr.load(it_local)
            if cand.get("parallel_layout") in (None, _layout_tag()):
                local_sd = cand
            elif not dist.is_initialized() or dist.get_rank() == 0:
                print(f"WARNING: local checkpoint of iteration {it_local} was written under layout {cand.get('parallel_layout')}; "
                      f"current layout {_layout_tag()}: ignoring it", flush=True)
    if local_sd is None and (not load_dir or not os.path.isfile(tracker)):
        if getattr(args, "exit_on_missing_checkpoint", False):
            raise SystemExit(f">> no checkpoint found in {load_dir}; exiting")
        return 0, 0
    if local_sd is not None:
        iteration, release, sd, is_sharded, base = int(local_sd["iteration"]), False, local_sd, False, None
        load_dir = f"{args.non_persistent_local_ckpt_dir} (local)"
        for i, m in enumerate(model):
            _unwrap(m).load_state_dict(sd["model" if len(model) == 1 else f"model{i}"], strict=strict)
        if optimizer is not None and not args.no_load_optim and "optimizer" in sd:
            optimizer.load_state_dict(sd["optimizer"])
        elif optimizer is not None:
            optimizer.reload_model_params()
    else:
        iteration, release = read_metadata(tracker)
        if getattr(args, "ckpt_step", None):
            iteration = args.ckpt_step
        base = get_checkpoint_name(load_dir, iteration, release, return_base_dir=True)
        is_sharded = os.path.exists(os.path.join(base, "metadata.json"))
    if local_sd is not None:
        pass
    elif is_sharded and torch_dist.is_torch_dist_checkpoint(base):
        # reference layout (.metadata + __*.distcp): model AND optimizer state reshard to the current layout
        sd = torch_dist.load_common_state_dict(base)
        want_opt = optimizer is not None and not args.finetune and not args.no_load_optim and "optimizer" in sd \
            and hasattr(optimizer, "load_sharded_common_state")
        with _full_params(model, writeback=True):
            ssd = _model_sharded_state_dict(model, args)
            md = torch_dist.read_metadata(base)
            if not any(k.startswith("decoder.layers.") and not k[len("decoder.layers."):].split(".")[0].isdigit()
                       for k in md.state_dict_metadata) and any(".layers." in k for k in md.state_dict_metadata):
                # checkpoint written with per-layer keys (non-homogeneous / MoE / older runs): ask for the same layout
                ssd = {}
                for i, m in enumerate(model):
                    ssd.update(_unwrap(m).sharded_state_dict(prefix="" if len(model) == 1 else f"model{i}."))
            model_keys = set(ssd)
            if want_opt:
                ssd.update(optimizer.sharded_state_dict(ssd, is_loading=True))
            mode = getattr(args, "dist_ckpt_strictness", None) or "assume_ok_unexpected"

            def _load(req):
                if mode == "assume_ok_unexpected":
                    return torch_dist.load(req, base, strict=strict)
                # --dist-ckpt-strictness: key mismatches are logged / raised / tolerated as StrictHandling says
                from ..ckpt import serialization

                res = serialization.load(req, base, strict=mode)
                if isinstance(res, tuple) and (res[1] or res[2]) and (not dist.is_initialized() or dist.get_rank() == 0):
                    print(f"checkpoint key mismatches: missing {res[1][:8]} unexpected {res[2][:8]}", flush=True)
                return res

            try:
                _load(ssd)
            except KeyError:
                if not want_opt:
                    raise
                # optimizer tensors absent (saved with --no-save-optim or by another optimizer): weights only
                _load({k: v for k, v in ssd.items() if k in model_keys})
                want_opt = False
        if optimizer is not None:
            rank = dist.get_rank() if dist.is_initialized() else 0
            op = os.path.join(base, f"optim_{rank:05d}.pt")
            if want_opt:
                optimizer.load_sharded_common_state(sd["optimizer"])
            elif (not args.finetune and not args.no_load_optim and os.path.exists(op)
                  and sd.get("parallel_layout") in (None, _layout_tag())):
                optimizer.load_state_dict(torch.load(op, map_location="cpu", weights_only=False))
            else:
                optimizer.reload_model_params()
        _fix_qkv_ordering_if_needed(model, sd)
    elif is_sharded:
        ssd = {}
        with _full_params(model, writeback=True):
            for i, m in enumerate(model):
                prefix = "" if len(model) == 1 else f"model{i}."
                ssd.update(_unwrap(m).sharded_state_dict(prefix=prefix))
            sharded.load(ssd, base, strict=strict)
        sd = sharded.load_common_state_dict(base)
        rank = dist.get_rank() if dist.is_initialized() else 0
        op = os.path.join(base, f"optim_{rank:05d}.pt")
        same_layout = sd.get("parallel_layout") in (None, _layout_tag())
        if optimizer is not None and not args.finetune and not args.no_load_optim and os.path.exists(op) and same_layout:
            optimizer.load_state_dict(torch.load(op, map_location="cpu", weights_only=False))
        elif optimizer is not None:
            if not same_layout and not args.finetune and not args.no_load_optim:
                print(f"WARNING: per-rank optimizer files were written under layout {sd.get('parallel_layout')}, current "
                      f"layout is {_layout_tag()}: optimizer state is re-initialised from the loaded weights", flush=True)
            optimizer.reload_model_params()
    else:
        name = get_checkpoint_name(load_dir, iteration, release)
        sd = torch.load(name, map_location="cpu", weights_only=False)
        for i, m in enumerate(model):
            target = _fsdp_wrapper(m) or _unwrap(m)          # ZeRO-3: load through the wrapper (gathers, loads, re-shards)
            msd = sd["model" if len(model) == 1 else f"model{i}"]
            if not hasattr(_unwrap(m), "kd_projections"):    # a distillation run's trainable projections are not model state
                msd = {k: v for k, v in msd.items() if not k.startswith("kd_projections.")}
            target.load_state_dict(msd, strict=strict)
        if optimizer is not None and not args.finetune and not args.no_load_optim:
            dp_rank = ps.get_data_parallel_rank(with_context_parallel=True)
            dop = os.path.join(os.path.dirname(name), "distrib_optim.pt" if dp_rank == 0 else f"distrib_optim_dp{dp_rank:03d}.pt")
            if "optimizer" in sd:
                optimizer.load_state_dict(sd["optimizer"])
            elif os.path.exists(dop):
                optimizer.load_state_dict(torch.load(dop, map_location="cpu", weights_only=False))
            else:
                optimizer.reload_model_params()
        elif optimizer is not None:
            optimizer.reload_model_params()
    if opt_param_scheduler is not None and sd.get("opt_param_scheduler") and not args.finetune:
        opt_param_scheduler.load_state_dict(sd["opt_param_scheduler"])
    if not release and not args.finetune and not args.no_load_rng and sd.get("rng_state"):
        try:
            st = sd["rng_state"][0]
            random.setstate(st["random_rng_state"])
            np.random.set_state(st["np_rng_state"])
            torch.set_rng_state(st["torch_rng_state"])
            if torch.cuda.is_available() and st.get("cuda_rng_state") is not None:
                torch.cuda.set_rng_state(st["cuda_rng_state"])
            if st.get("rng_tracker_states"):
                get_cuda_rng_tracker().set_states(st["rng_tracker_states"])
        except Exception as e:  # noqa: BLE001
            print(f"WARNING: could not restore rng state: {e}", flush=True)
    if args.finetune or release:
        iteration = 0
    else:
        ca = sd.get("args")
        if ca is not None:
            args.consumed_train_samples = getattr(ca, "consumed_train_samples", 0)
            args.consumed_valid_samples = getattr(ca, "consumed_valid_samples", 0)
    if dist.is_initialized():
        dist.barrier()
    if load_arg == "load":
        from . import ft_integration

        ft_integration.on_checkpoint_loaded(is_local_chkpt=local_sd is not None)
    print(f"  successfully loaded checkpoint from {load_dir} at iteration {iteration}", flush=True) \
        if (not dist.is_initialized() or dist.get_rank() == 0) else None
    return iteration, sd.get("num_floating_point_operations_so_far", 0)


def _fix_qkv_ordering_if_needed(model, sd) -> None:
    """Checkpoints older than version 2.0 store the fused QKV projection as [q_all; k_all; v_all] (version 0) or
    [heads, 3, hn] (version 1); current layout is per-group [q..., k, v] (reference checkpointing.py:822
    ``fix_query_key_value_ordering``)."""
    version = sd.get("checkpoint_version", _CHECKPOINT_VERSION)
    if version is None or version >= 2.0:
        return
    for m in model:
        core = _unwrap(m)
        for name, p in core.named_parameters():
            if not name.endswith(("linear_qkv.weight", "linear_qkv.bias", "query_key_value.weight", "query_key_value.bias")):
                continue
            cfg = core.config
            tp = ps.get_tensor_model_parallel_world_size()
            heads = cfg.num_attention_heads // tp
            hn = cfg.kv_channels or cfg.hidden_size // cfg.num_attention_heads
            rest = tuple(p.shape[1:])
            with torch.no_grad():
                if version == 0:
                    t = p.data.view(3, heads, hn, *rest).transpose(0, 1).contiguous()
                else:
                    t = p.data.view(heads, hn, 3, *rest).transpose(1, 2).contiguous()
                p.data.copy_(t.view(p.shape))


def load_args_from_checkpoint(args, load_arg="load"):
    load_dir = getattr(args, load_arg)
    tracker = get_checkpoint_tracker_filename(load_dir)
    if not os.path.isfile(tracker):
        return args
    iteration, release = read_metadata(tracker)
    base = get_checkpoint_name(load_dir, iteration, release, return_base_dir=True)
    if os.path.exists(os.path.join(base, "common.pt")):
        sd = sharded.load_common_state_dict(base)     # same file name / format in both sharded layouts
    else:
        name = get_checkpoint_name(load_dir, iteration, release, tensor_rank=0, pipeline_rank=0,
                                   pipeline_parallel=os.path.isdir(os.path.join(base, "mp_rank_00_000")),
                                   expert_parallel=False, expert_rank=0)
        sd = torch.load(name, map_location="cpu", weights_only=False)
    ca = sd.get("args")
    if ca is None:
        return args
    for k in ("num_layers", "hidden_size", "ffn_hidden_size", "seq_length", "num_attention_heads", "num_query_groups",
              "group_query_attention", "kv_channels", "max_position_embeddings", "position_embedding_type",
              "add_position_embedding", "use_rotary_position_embeddings", "rotary_base", "rotary_percent",
              "add_bias_linear", "add_qkv_bias", "swiglu", "untie_embeddings_and_output_weights",
              "apply_layernorm_1p", "normalization", "tokenizer_type", "padded_vocab_size", "apply_query_key_layer_scaling",
              "num_experts", "moe_router_topk"):
        if hasattr(ca, k):
            setattr(args, k, getattr(ca, k))
    return args


# -- further names of the reference's ``megatron/training/checkpointing.py`` --------------------------------------------------
from enum import Enum, auto as _auto  # noqa: E402

_LOADED_CHECKPOINT_VERSION = None


class CheckpointType(Enum):
    LEGACY = _auto()
    LOCAL = _auto()
    GLOBAL = _auto()
    TORCH_DCP = _auto()


def set_checkpoint_version(value):
    global _LOADED_CHECKPOINT_VERSION
    if _LOADED_CHECKPOINT_VERSION is not None and _LOADED_CHECKPOINT_VERSION != value:
        raise AssertionError("checkpoint versions do not match")
    _LOADED_CHECKPOINT_VERSION = value


def get_checkpoint_version():
    return _LOADED_CHECKPOINT_VERSION


def ensure_directory_exists(filename, check_parent: bool = True):
    os.makedirs(os.path.dirname(filename) if check_parent else filename, exist_ok=True)


def checkpoint_exists(checkpoints_path) -> bool:
    """A tracker file (``latest_checkpointed_iteration.txt``) marks a directory that holds at least one finished checkpoint."""
    if checkpoints_path is None:
        return False
    return os.path.isfile(get_checkpoint_tracker_filename(checkpoints_path)) or \
        os.path.isfile(os.path.join(checkpoints_path, "latest"))


def get_distributed_optimizer_checkpoint_name(model_checkpoint_name):
    return os.path.join(os.path.dirname(model_checkpoint_name), "distrib_optim.pt")


def find_checkpoint_rank_0(checkpoints_path, iteration, release=False):
    """File (legacy layouts, with / without pipeline / expert parallel directory names) or directory (distributed formats) of
    rank 0's part of a checkpoint; ``None`` when there is none."""
    for pp in (False, True):
        for ep in (False, True):
            name = get_checkpoint_name(checkpoints_path, iteration, release, pipeline_parallel=pp, tensor_rank=0, pipeline_rank=0,
                                       expert_parallel=ep, expert_rank=0)
            if os.path.isfile(name):
                return name
    base = get_checkpoint_name(checkpoints_path, iteration, release, return_base_dir=True)
    if os.path.exists(os.path.join(base, "metadata.json")) or os.path.exists(os.path.join(base, ".metadata")):
        return base
    return None


def fix_query_key_value_ordering(model, checkpoint_version):
    """Checkpoints older than version 2.0 store the fused QKV weight in another head order; current ones need no change."""
    if checkpoint_version is not None and checkpoint_version < 2.0:
        _fix_qkv_ordering_if_needed(model if isinstance(model, (list, tuple)) else [model], {"checkpoint_version": checkpoint_version})


def cleanup_old_non_persistent_checkpoint(save_dir, leave_ckpt_num: int = 1, do_async: bool = False):
    """Keep the newest ``leave_ckpt_num`` ``iter_*`` directories of a non-persistent checkpoint folder."""
    import shutil

    if (dist.is_initialized() and dist.get_rank() != 0) or not os.path.isdir(save_dir):
        return
    dirs = sorted(d for d in os.listdir(save_dir) if d.startswith("iter_"))
    for d in dirs[:-leave_ckpt_num] if leave_ckpt_num > 0 else dirs:
        shutil.rmtree(os.path.join(save_dir, d), ignore_errors=True)
