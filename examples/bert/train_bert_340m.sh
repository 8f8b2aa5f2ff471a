#!/bin/bash
source "$(dirname "$0")/../_common.sh"
$LAUNCH $ROOT/pretrain_bert.py \
  --tensor-model-parallel-size ${TP:-1} --num-layers 24 --hidden-size 1024 --num-attention-heads 16 \
  --seq-length 512 --max-position-embeddings 512 --micro-batch-size 8 --global-batch-size 256 \
  --train-iters ${TRAIN_ITERS:-1000000} --lr 1e-4 --min-lr 1e-5 --lr-decay-style linear --weight-decay 1e-2 \
  --clip-grad 1.0 --bf16 --tokenizer-type ${TOKENIZER:-NullTokenizer} --vocab-size ${VOCAB_SIZE:-30592} \
  --log-interval 10 --save $CHECKPOINT_PATH --load $CHECKPOINT_PATH "$@"
