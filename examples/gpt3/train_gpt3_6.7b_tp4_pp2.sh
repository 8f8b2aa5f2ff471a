#!/bin/bash
# GPT-3 6.7B on one 8xB200 node: TP4 x PP2 + sequence parallel, fused AG->GEMM over the NVLink arena
# (BASELINE.json config #2; reference examples/gpt3/train_gpt3_175b_distributed.sh).
source "$(dirname "$0")/../_common.sh"
$LAUNCH $ROOT/pretrain_gpt.py \
  --tensor-model-parallel-size 4 --pipeline-model-parallel-size 2 --sequence-parallel \
  --num-layers 32 --hidden-size 4096 --num-attention-heads 32 --seq-length 2048 --max-position-embeddings 2048 \
  --micro-batch-size 4 --global-batch-size 64 --train-iters ${TRAIN_ITERS:-500000} \
  --lr 1.2e-4 --min-lr 1.2e-5 --lr-decay-style cosine --lr-warmup-fraction 0.001 --weight-decay 0.1 \
  --adam-beta1 0.9 --adam-beta2 0.95 --clip-grad 1.0 --init-method-std 0.006 --bf16 \
  --attention-dropout 0.0 --hidden-dropout 0.0 \
  --log-interval 10 --save-interval 1000 --eval-interval 1000 --eval-iters 10 --log-throughput \
  --save $CHECKPOINT_PATH --load $CHECKPOINT_PATH --tensorboard-dir $TENSORBOARD_DIR $DATA_ARGS "$@"
