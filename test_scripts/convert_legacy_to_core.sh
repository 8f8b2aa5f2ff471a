#!/bin/bash
# legacy (--use-legacy-models) checkpoint -> core layout, optionally re-sharded
source "$(dirname "$0")/_env.sh"
python $ROOT/tools/checkpoint/convert.py --model-type GPT --loader legacy --saver core --load-dir ${LOAD_DIR:?set LOAD_DIR} \
  --save-dir ${SAVE_DIR:-$WORK/core} --target-tensor-parallel-size ${TP:-1} --target-pipeline-parallel-size ${PP:-1} "$@"
