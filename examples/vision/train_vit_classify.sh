#!/bin/bash
source "$(dirname "$0")/../_common.sh"
$LAUNCH $ROOT/pretrain_vision_classify.py \
  --tensor-model-parallel-size ${TP:-1} --num-layers 12 --hidden-size 768 --num-attention-heads 12 --seq-length 197 \
  --max-position-embeddings 197 --img-h 224 --img-w 224 --patch-dim 16 --num-classes 1000 --micro-batch-size 64 \
  --global-batch-size 1024 --train-iters ${TRAIN_ITERS:-100000} --lr 5e-4 --min-lr 1e-5 --lr-decay-style cosine \
  --weight-decay 0.05 --clip-grad 1.0 --bf16 --mock-data --tokenizer-type NullTokenizer --vocab-size 8 \
  --log-interval 10 --save $CHECKPOINT_PATH "$@"
