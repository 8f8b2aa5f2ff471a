#!/bin/bash
# Continuous batching with a paged KV cache (40 GB of KV blocks of 256 tokens) behind the REST / OpenAI-style endpoints.
source "$(dirname "$0")/../_common.sh"
$LAUNCH $ROOT/tools/run_text_generation_server.py \
  --tensor-model-parallel-size ${TP:-$GPUS_PER_NODE} --num-layers 32 --hidden-size 4096 --num-attention-heads 32 \
  --seq-length 4096 --max-position-embeddings 4096 --bf16 --tokenizer-type ${TOKENIZER:-NullTokenizer} \
  --vocab-size ${VOCAB_SIZE:-50257} --load $CHECKPOINT_PATH --port ${PORT:-5000} --inference-max-seq-length 4096 \
  --inference-dynamic-batching --inference-dynamic-batching-buffer-size-gb 40 --inference-dynamic-batching-chunk-size 256 "$@"
