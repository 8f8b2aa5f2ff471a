#!/bin/bash
# Shared launcher bits: one process per GPU on this node, NCCL over NVLink, rendezvous on loopback.
GPUS_PER_NODE=${GPUS_PER_NODE:-8}
MASTER_ADDR=${MASTER_ADDR:-127.0.0.1}
MASTER_PORT=${MASTER_PORT:-6000}
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
LAUNCH="python -m torch.distributed.run --nnodes=1 --nproc-per-node $GPUS_PER_NODE --master-addr $MASTER_ADDR --master-port $MASTER_PORT"
CHECKPOINT_PATH=${CHECKPOINT_PATH:-$ROOT/checkpoints}
TENSORBOARD_DIR=${TENSORBOARD_DIR:-$ROOT/tensorboard}
# data: pass VOCAB_FILE / MERGE_FILE / DATA_PATH for a real corpus, otherwise synthetic tokens
if [ -n "$DATA_PATH" ]; then
  DATA_ARGS="--data-path $DATA_PATH --vocab-file $VOCAB_FILE --merge-file $MERGE_FILE --split 949,50,1"
else
  DATA_ARGS="--mock-data --tokenizer-type NullTokenizer --vocab-size ${VOCAB_SIZE:-50257}"
fi
