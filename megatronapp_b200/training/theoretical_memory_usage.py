"""Closed-form estimate of per-GPU weight/optimizer and activation memory for a run.

Parity: reference ``training/theoretical_memory_usage.py`` (``report_theoretical_memory``; invoked by
``--report-theoretical-memory`` style tooling, ``tools/report_theoretical_memory.py``).  Formulas
follow the activation-recomputation analysis (34·s·b·h + 5·a·s²·b bytes per layer without
recompute; sequence parallelism divides everything by tp; selective recompute drops the 5·a·s²·b
term -- which flash attention removes anyway in this framework) and Megatron's 1F1B in-flight
microbatch count.  Sized against 180 GB of HBM3e per B200.
"""
from __future__ import annotations

NUM_BYTES_IN_MEGABYTE = 1024 * 1024
B200_HBM_BYTES = 180e9


def _padded_vocab(args):
    v = getattr(args, "padded_vocab_size", None)
    if v:
        return v
    mult = getattr(args, "make_vocab_size_divisible_by", 128) * args.tensor_model_parallel_size
    raw = getattr(args, "vocab_size", None) or 50257
    return (raw + mult - 1) // mult * mult


def compute_weight_and_optimizer_memory(args, verbose=False):
    h, L = args.hidden_size, args.num_layers
    V = _padded_vocab(args)
    kv = getattr(args, "kv_channels", None) or h // args.num_attention_heads
    nq = args.num_attention_heads
    ng = getattr(args, "num_query_groups", None) or nq
    ffn = args.ffn_hidden_size
    glu = 2 if getattr(args, "swiglu", False) else 1
    n_exp = getattr(args, "num_experts", None) or 1
    attn = h * (nq * kv) + 2 * h * (ng * kv) + (nq * kv) * h          # qkv + proj
    mlp = n_exp * (glu * h * ffn + ffn * h) + (h * n_exp if n_exp > 1 else 0)
    norms = 2 * 2 * h
    per_layer = attn + mlp + norms
    emb = V * h
    untied = getattr(args, "untie_embeddings_and_output_weights", False)
    total = L * per_layer + emb * (2 if untied else 1) + 2 * h
    tp, pp = args.tensor_model_parallel_size, args.pipeline_model_parallel_size
    # most loaded shard: first stage holds the embedding (and, tied & pp==1, nothing else extra)
    on_first = (L // pp) * per_layer + emb
    if untied and pp == 1:
        on_first += emb
    on_shard = on_first / tp
    dp = max(1, getattr(args, "data_parallel_size", 1))
    # bf16 param (2) + fp32 main grad (4) + fp32 master/m/v (12); ZeRO-1 shards the 12 over dp
    bytes_per_param = 18 if not getattr(args, "use_distributed_optimizer", False) else 6 + 12 / dp
    mem = on_shard * bytes_per_param
    if verbose:
        print(f"Number of parameters in transformer layers in billions: {L * per_layer / 1e9:.2f}")
        print(f"Number of parameters in embedding layers in billions: {emb * (2 if untied else 1) / 1e9:.2f}")
        print(f"Total number of parameters in billions: {total / 1e9:.2f}")
        print(f"Number of parameters in most loaded shard in billions: {on_shard / 1e9:.4f}")
    return mem


def compute_activation_memory(args, num_microbatches=None, verbose=False):
    s, b, h = args.seq_length, args.micro_batch_size, args.hidden_size
    a = args.num_attention_heads
    tp, pp = args.tensor_model_parallel_size, args.pipeline_model_parallel_size
    L = args.num_layers
    ffn_ratio = args.ffn_hidden_size / h
    # per layer, bytes (bf16 activations, 1-byte dropout masks); flash attention keeps no s^2 tensor
    per_layer = s * b * h * (10 + 2 * 2 * ffn_ratio + 4 + 2)      # attn block + mlp block + norms/dropout
    sp = getattr(args, "sequence_parallel", False)
    if sp:
        per_layer = per_layer / tp
    elif tp > 1:
        # without SP only the tensor-parallel regions (qkv/attn/fc1/act) shrink; norms, dropouts and
        # the block inputs stay replicated
        replicated = s * b * h * 10
        per_layer = replicated + (per_layer - replicated) / tp
    if getattr(args, "recompute_granularity", None) == "full":
        per_layer = 2 * s * b * h / (tp if sp else 1)
    mem = per_layer * (L // pp)
    # embedding dropout + output layer (logits fp32 for the loss on the last stage)
    mem += 2 * s * b * h / (tp if sp else 1)
    vpp = getattr(args, "virtual_pipeline_model_parallel_size", None)
    in_flight = pp
    if vpp:
        in_flight = pp * (1 + (pp - 1) / (pp * vpp))
    if num_microbatches is not None and num_microbatches < pp:
        in_flight = num_microbatches
    mem *= in_flight if pp > 1 else 1
    if pp == 1:
        mem += s * b * _padded_vocab(args) / tp * (2 + 4)
    if verbose:
        print(f"Activation memory footprint per transformer layer: {per_layer / NUM_BYTES_IN_MEGABYTE / 1:.1f} MB")
    return mem


def report_theoretical_memory(args, num_microbatches=None, verbose=False):
    w = compute_weight_and_optimizer_memory(args, verbose)
    a = compute_activation_memory(args, num_microbatches, verbose)
    total = w + a
    print(f"Theoretical memory footprints: weight and optimizer={w / NUM_BYTES_IN_MEGABYTE:.2f} MB, "
          f"activation={a / NUM_BYTES_IN_MEGABYTE:.2f} MB, total={total / NUM_BYTES_IN_MEGABYTE:.2f} MB "
          f"({100.0 * total / B200_HBM_BYTES:.1f}% of one B200's 180 GB)")
    return w, a, total
