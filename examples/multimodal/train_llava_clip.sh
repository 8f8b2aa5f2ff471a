#!/bin/bash
source "$(dirname "$0")/../_common.sh"
$LAUNCH $ROOT/pretrain_vlm.py \
  --tensor-model-parallel-size ${TP:-1} --num-layers 32 --hidden-size 4096 --num-attention-heads 32 \
  --seq-length 1024 --max-position-embeddings 4096 --micro-batch-size 2 --global-batch-size 64 \
  --img-h 336 --img-w 336 --patch-dim 14 --vision-num-layers 24 --vision-hidden-size 1024 --vision-num-attention-heads 16 \
  --disable-vision-class-token --freeze-LM --normalization RMSNorm --swiglu --position-embedding-type rope \
  --disable-bias-linear --untie-embeddings-and-output-weights --tokenizer-type NullTokenizer --vocab-size 32000 \
  --train-iters ${TRAIN_ITERS:-20000} --lr 1e-3 --lr-decay-style cosine --bf16 --log-interval 10 \
  --save $CHECKPOINT_PATH --load $CHECKPOINT_PATH "$@"
