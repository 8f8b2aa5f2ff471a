#!/bin/bash
# MegaFBD: forward and backward of each data-parallel pair run on separate GPUs
# (reference root pretrain_gpt.sh with --forward-backward-disaggregating).
source "$(dirname "$0")/../_common.sh"
$LAUNCH $ROOT/pretrain_gpt.py --pipeline-model-parallel-size 2 \
  --num-layers 24 --hidden-size 2048 --num-attention-heads 16 --seq-length 2048 --max-position-embeddings 2048 \
  --micro-batch-size 2 --global-batch-size 32 --train-iters ${TRAIN_ITERS:-1000} --lr 1e-4 --bf16 --log-interval 10 \
  --forward-backward-disaggregating --ignore-forward-tensor-parallel $DATA_ARGS "$@"
