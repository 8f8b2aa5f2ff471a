#!/bin/bash
# Mamba-2 / attention / MLP hybrid (pattern via --hybrid-attention-ratio / --hybrid-mlp-ratio).
source "$(dirname "$0")/../_common.sh"
$LAUNCH $ROOT/pretrain_mamba.py \
  --tensor-model-parallel-size ${TP:-1} --num-layers 56 --hidden-size 4096 --num-attention-heads 32 \
  --group-query-attention --num-query-groups 8 --hybrid-attention-ratio 0.08 --hybrid-mlp-ratio 0.5 \
  --seq-length 4096 --max-position-embeddings 4096 --micro-batch-size 1 --global-batch-size 64 \
  --normalization RMSNorm --position-embedding-type none --disable-bias-linear --untie-embeddings-and-output-weights \
  --train-iters ${TRAIN_ITERS:-100000} --lr 2.5e-4 --min-lr 2.5e-5 --lr-decay-style cosine --bf16 \
  --log-interval 10 --save $CHECKPOINT_PATH --load $CHECKPOINT_PATH $DATA_ARGS "$@"
