#!/bin/bash
# MegaDPP: dynamic pipeline scheduling over the NVLink mailbox (reference root run_single_gpt.sh).
# GPT-3 1.3B, PP4 x DP2, 4 model chunks per stage (BASELINE.json config #4).
source "$(dirname "$0")/../_common.sh"
$LAUNCH $ROOT/pretrain_gpt.py --pipeline-model-parallel-size 4 --num-layers-per-virtual-pipeline-stage 2 \
  --num-layers 24 --hidden-size 2048 --num-attention-heads 16 --seq-length 2048 --max-position-embeddings 2048 \
  --micro-batch-size 2 --global-batch-size 32 --train-iters ${TRAIN_ITERS:-1000} --lr 1e-4 --bf16 --log-interval 10 \
  --use-dpp --dpp-policy ${DPP_POLICY:-auto} $DATA_ARGS "$@"
