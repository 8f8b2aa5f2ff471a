"""Legacy (vision / retrieval) data pipelines of the reference's ``megatron/legacy/data``."""
