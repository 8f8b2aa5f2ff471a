#!/bin/bash
# MegaScan: trace a short run (TP 2), then aggregate to a Chrome trace and run the slow-rank detector
source "$(dirname "$0")/_env.sh"
[ -f $DATA_PREFIX.idx ] || bash $ROOT/test_scripts/process_dataset_gpt.sh
TRACE_DIR=${TRACE_DIR:-$WORK/trace}
rm -rf $TRACE_DIR
$LAUNCH $ROOT/pretrain_gpt.py $MODEL_ARGS $TRAIN_ARGS $TOK_ARGS --data-path $DATA_PREFIX --tensor-model-parallel-size 2 \
  --sequence-parallel --trace --trace-dir $TRACE_DIR --trace-interval 5 --continuous-trace-iterations 2 --trace-granularity full "$@" || exit 1
bash $ROOT/test_scripts/aggregate_trace.sh $TRACE_DIR
