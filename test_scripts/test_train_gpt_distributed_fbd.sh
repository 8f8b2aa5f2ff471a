#!/bin/bash
# MegaFBD: forward and backward of every stage on separate ranks (world = 2 x the usual layout)
source "$(dirname "$0")/_env.sh"
[ -f $DATA_PREFIX.idx ] || bash $ROOT/test_scripts/process_dataset_gpt.sh
$LAUNCH $ROOT/pretrain_gpt.py $MODEL_ARGS $TRAIN_ARGS $TOK_ARGS --data-path $DATA_PREFIX --forward-backward-disaggregating "$@"
