#!/bin/bash
# Llama-3 8B, TP8, fp8 (hybrid E4M3/E5M2, tensor-wise scaling) with MegaScan tracing
# (BASELINE.json config #5).  Throttle one GPU with scripts/gpu_control.sh (root) or --slow-rank-inject
# and run scripts/aggregate.py -d to see the detector point at it.
source "$(dirname "$0")/../_common.sh"
$LAUNCH $ROOT/pretrain_gpt.py \
  --tensor-model-parallel-size 8 --sequence-parallel \
  --num-layers 32 --hidden-size 4096 --ffn-hidden-size 14336 --num-attention-heads 32 \
  --group-query-attention --num-query-groups 8 --seq-length 8192 --max-position-embeddings 8192 \
  --normalization RMSNorm --swiglu --position-embedding-type rope --rotary-base 500000 \
  --disable-bias-linear --untie-embeddings-and-output-weights \
  --fp8-format hybrid --fp8-recipe tensorwise \
  --micro-batch-size 1 --global-batch-size 64 --train-iters ${TRAIN_ITERS:-100000} \
  --lr 3e-4 --min-lr 3e-5 --lr-decay-style cosine --weight-decay 0.1 --clip-grad 1.0 --bf16 \
  --trace --trace-dir ${TRACE_DIR:-trace} --trace-interval 5 --continuous-trace-iterations 2 \
  --log-interval 10 --save $CHECKPOINT_PATH --load $CHECKPOINT_PATH $DATA_ARGS "$@"
