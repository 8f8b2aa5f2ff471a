#!/bin/bash
# GPT-3 1.3B, pure data parallel with the ZeRO-1 distributed optimizer and the fused scale/cast
# gradient reduce-scatter overlapped with backward.
source "$(dirname "$0")/../_common.sh"
$LAUNCH $ROOT/pretrain_gpt.py \
  --num-layers 24 --hidden-size 2048 --num-attention-heads 16 --seq-length 2048 --max-position-embeddings 2048 \
  --micro-batch-size 4 --global-batch-size 256 --train-iters ${TRAIN_ITERS:-300000} \
  --use-distributed-optimizer --overlap-grad-reduce --overlap-param-gather --fused-grad-reduce-scatter \
  --lr 2.0e-4 --min-lr 2.0e-5 --lr-decay-style cosine --weight-decay 0.1 --clip-grad 1.0 --bf16 \
  --log-interval 10 --save-interval 2000 --eval-interval 1000 --eval-iters 10 \
  --save $CHECKPOINT_PATH --load $CHECKPOINT_PATH $DATA_ARGS "$@"
