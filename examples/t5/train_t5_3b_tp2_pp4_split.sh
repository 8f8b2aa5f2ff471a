#!/bin/bash
# T5 with the encoder/decoder pipeline split: stages 0-1 hold the encoder, stages 2-3 the decoder (tp2 x pp4 = 8 GPUs).
source "$(dirname "$0")/../_common.sh"
$LAUNCH $ROOT/pretrain_t5.py \
  --tensor-model-parallel-size 2 --pipeline-model-parallel-size 4 --pipeline-model-parallel-split-rank 2 \
  --encoder-num-layers 24 --decoder-num-layers 24 --hidden-size 1024 --num-attention-heads 32 --kv-channels 128 \
  --ffn-hidden-size 16384 --encoder-seq-length 512 --decoder-seq-length 128 --max-position-embeddings 512 \
  --position-embedding-type relative --relative-attention-num-buckets 32 --relative-attention-max-distance 128 \
  --micro-batch-size 8 --global-batch-size 256 --train-iters ${TRAIN_ITERS:-500000} --lr 1e-4 --min-lr 1e-5 \
  --lr-decay-style linear --weight-decay 1e-2 --clip-grad 1.0 --bf16 --mock-data --tokenizer-type NullTokenizer \
  --vocab-size ${VOCAB_SIZE:-32128} --log-interval 10 --save $CHECKPOINT_PATH --load $CHECKPOINT_PATH "$@"
