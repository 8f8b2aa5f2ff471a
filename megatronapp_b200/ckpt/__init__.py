"""Distributed checkpoints: ``sharded`` (own format), ``torch_dist`` (the reference's PyTorch-DCP layout) and the
``dist_checkpointing``-style entry points in ``serialization`` (strict modes, metadata / plain loads, key removal, sharding
validation)."""
from .serialization import (CheckpointingException, LocalNonpersistentObject, StrictHandling,  # noqa: F401
                            check_is_distributed_checkpoint, load, load_common_state_dict, load_plain_tensors,
                            load_sharded_metadata, load_tensors_metadata, remove_sharded_tensors, save,
                            validate_sharding_integrity)
from .sharded import ShardedObject, ShardedTensor  # noqa: F401
