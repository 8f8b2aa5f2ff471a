#!/bin/bash
# Knowledge distillation: a 1.3B student learns from a frozen 6.7B teacher checkpoint (logit KL + one intermediate pair).
#   TEACHER_PATH/model_config.yaml may describe the teacher (num_layers, hidden_size, num_attention_heads, ...); without it the
#   architecture recorded in the checkpoint is used.  KD_CFG is the distillation yaml (logit_layers, intermediate_layer_pairs,
#   skip_lm_loss, kd_loss_scale); unset = logits only.  Afterwards: add --export-kd-finalize --load $CHECKPOINT_PATH to save the
#   plain student.
source "$(dirname "$0")/../_common.sh"
TEACHER_PATH=${TEACHER_PATH:?checkpoint directory of the teacher}
$LAUNCH $ROOT/pretrain_gpt.py \
  --tensor-model-parallel-size ${TP:-2} --sequence-parallel --num-layers 24 --hidden-size 2048 --num-attention-heads 16 \
  --seq-length 2048 --max-position-embeddings 2048 --micro-batch-size 4 --global-batch-size 64 --train-iters ${TRAIN_ITERS:-1000} \
  --lr 1.0e-4 --min-lr 1.0e-5 --lr-decay-style cosine --lr-warmup-iters 50 --bf16 --use-distributed-optimizer --overlap-grad-reduce \
  --export-kd-teacher-load $TEACHER_PATH ${KD_CFG:+--export-kd-cfg $KD_CFG} \
  $DATA_ARGS --save $CHECKPOINT_PATH --save-interval 500 --log-interval 10 --eval-interval 500 --eval-iters 10 "$@"
