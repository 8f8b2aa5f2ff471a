#!/bin/bash
# per-rank benchmark-data-*.json -> one Chrome trace (open in chrome://tracing / Perfetto) + abnormal.txt when a rank is slow
source "$(dirname "$0")/_env.sh"
TRACE_DIR=${1:-$WORK/trace}
python $ROOT/scripts/aggregate.py -b $TRACE_DIR -o $TRACE_DIR/aggregated.json -d --abnormal-file $TRACE_DIR/abnormal.txt
ls -l $TRACE_DIR/aggregated.json
[ -f $TRACE_DIR/abnormal.txt ] && echo "abnormal rank: $(cat $TRACE_DIR/abnormal.txt)" || echo "no abnormal rank detected"
