"""Function-level API of ``megatron.core`` that user scripts import directly (``megatron.core.utils``, ``...enums``,
``...optimizer.clip_grads``, ``...tensor_parallel.utils``, ``...transformer.utils``, ``...fusions.*``, the Transformer-Engine
layer names, the legacy ``text_generation`` helpers), implemented on this framework's kernels and modules.  ``compat/megatron``
maps the reference's module paths onto these files."""
