"""Distributed checkpoints: ``sharded`` (own format), ``torch_dist`` (the reference's PyTorch-DCP layout) and the
``dist_checkpointing``-style entry points in ``serialization`` (strict modes, metadata / plain loads, key removal, sharding
validation).  Names resolve lazily: ``torch.distributed.checkpoint`` is only imported when one of them is used."""
_SERIALIZATION = ("CheckpointingException", "LocalNonpersistentObject", "StrictHandling", "check_is_distributed_checkpoint", "load",
                  "load_common_state_dict", "load_plain_tensors", "load_sharded_metadata", "load_tensors_metadata",
                  "remove_sharded_tensors", "save", "validate_sharding_integrity")
_SHARDED = ("ShardedObject", "ShardedTensor")


def __getattr__(name):
    import importlib

    if name in _SERIALIZATION:
        return getattr(importlib.import_module(".serialization", __name__), name)
    if name in _SHARDED:
        return getattr(importlib.import_module(".sharded", __name__), name)
    if name in ("serialization", "sharded", "torch_dist"):
        return importlib.import_module("." + name, __name__)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")


def __dir__():
    return sorted(list(globals()) + list(_SERIALIZATION) + list(_SHARDED))
