#!/bin/bash
# MegaScan: trace a short run, inject a slow rank in software, aggregate to a Chrome trace and run the detector.
source "$(dirname "$0")/../_common.sh"
TRACE_DIR=${TRACE_DIR:-trace}
$LAUNCH $ROOT/pretrain_gpt.py --tensor-model-parallel-size 2 --pipeline-model-parallel-size 2 \
  --num-layers 16 --hidden-size 2048 --num-attention-heads 16 --seq-length 2048 --max-position-embeddings 2048 \
  --micro-batch-size 2 --global-batch-size 32 --train-iters 20 --lr 1e-4 --bf16 --log-interval 5 \
  --trace --trace-dir $TRACE_DIR --trace-interval 5 --continuous-trace-iterations 2 --trace-granularity full \
  --slow-rank-inject ${SLOW_RANK:-3}:${SLOW_CYCLES:-200000000} $DATA_ARGS "$@"
python $ROOT/scripts/aggregate.py -b $TRACE_DIR -o $TRACE_DIR/aggregated.json -d
