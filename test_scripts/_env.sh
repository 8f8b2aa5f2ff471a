#!/bin/bash
# Common settings of the add-on smoke tests (reference ``test_scripts/``).  On a machine without GPUs the scripts fall back
# to a tiny model on CPU / gloo so every step of the workflow can still be exercised.
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
WORK=${WORK:-$ROOT/test_scripts/work}
mkdir -p "$WORK"
if [ -e /dev/nvidia0 ] && [ -z "$FORCE_CPU" ]; then
  NPROC=${NPROC:-4}
  BACKEND=nccl
  MODEL_ARGS=${MODEL_ARGS:-"--num-layers 12 --hidden-size 768 --num-attention-heads 12 --seq-length 1024 --max-position-embeddings 1024 --micro-batch-size 4 --global-batch-size 32 --bf16"}
else
  NPROC=${NPROC:-2}
  BACKEND=gloo
  export CUDA_VISIBLE_DEVICES=""
  MODEL_ARGS=${MODEL_ARGS:-"--num-layers 4 --hidden-size 64 --num-attention-heads 4 --seq-length 64 --max-position-embeddings 64 --micro-batch-size 2 --global-batch-size 8 --no-gradient-accumulation-fusion"}
fi
PORT=${MASTER_PORT:-$((20000 + RANDOM % 20000))}
LAUNCH="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NPROC --master-addr 127.0.0.1 --master-port $PORT"
TOK_ARGS="--tokenizer-type GPT2BPETokenizer --vocab-file $ROOT/datasets_gpt/vocab.json --merge-file $ROOT/datasets_gpt/merges.txt"
DATA_PREFIX=$WORK/gpt_text_document
TRAIN_ARGS="--train-iters ${TRAIN_ITERS:-10} --lr 1e-3 --log-interval 1 --eval-iters 1 --eval-interval 1000 --distributed-backend $BACKEND --split 98,1,1"
