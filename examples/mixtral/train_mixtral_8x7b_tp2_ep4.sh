#!/bin/bash
# Mixtral 8x7B, TP2 x EP4 on 8xB200 (BASELINE.json config #3): grouped expert GEMMs, dispatch over the
# NVLink arena ("flex" = peer-scatter kernel; "alltoall" = NCCL).
source "$(dirname "$0")/../_common.sh"
$LAUNCH $ROOT/pretrain_gpt.py \
  --tensor-model-parallel-size 2 --expert-model-parallel-size 4 --sequence-parallel \
  --num-layers 32 --hidden-size 4096 --ffn-hidden-size 14336 --num-attention-heads 32 \
  --group-query-attention --num-query-groups 8 --seq-length 4096 --max-position-embeddings 32768 \
  --num-experts 8 --moe-router-topk 2 --moe-router-load-balancing-type aux_loss --moe-aux-loss-coeff 1e-2 \
  --moe-grouped-gemm --moe-token-dispatcher-type ${DISPATCHER:-alltoall} \
  --normalization RMSNorm --swiglu --position-embedding-type rope --rotary-base 1000000 \
  --disable-bias-linear --untie-embeddings-and-output-weights --use-distributed-optimizer \
  --micro-batch-size 1 --global-batch-size 256 --train-iters ${TRAIN_ITERS:-500000} \
  --lr 1e-4 --min-lr 1e-5 --lr-decay-style cosine --weight-decay 0.1 --clip-grad 1.0 --bf16 \
  --log-interval 10 --save $CHECKPOINT_PATH --load $CHECKPOINT_PATH $DATA_ARGS "$@"
