#!/bin/bash
source "$(dirname "$0")/../_common.sh"
$LAUNCH $ROOT/pretrain_t5.py \
  --tensor-model-parallel-size ${TP:-1} --encoder-num-layers 12 --decoder-num-layers 12 --hidden-size 768 \
  --num-attention-heads 12 --kv-channels 64 --ffn-hidden-size 3072 --encoder-seq-length 512 --decoder-seq-length 128 \
  --max-position-embeddings 512 --micro-batch-size 16 --global-batch-size 512 --train-iters ${TRAIN_ITERS:-1000000} \
  --lr 1e-4 --min-lr 1e-5 --lr-decay-style linear --weight-decay 1e-2 --clip-grad 1.0 --bf16 \
  --tokenizer-type ${TOKENIZER:-NullTokenizer} --vocab-size ${VOCAB_SIZE:-30592} --vocab-extra-ids 100 \
  --log-interval 10 --save $CHECKPOINT_PATH --load $CHECKPOINT_PATH "$@"
