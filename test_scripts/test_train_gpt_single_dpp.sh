#!/bin/bash
# MegaDPP: interleaved pipeline (PP 2, 2 chunks per rank) with the dynamic scheduler over the NVLink mailbox
source "$(dirname "$0")/_env.sh"
[ -f $DATA_PREFIX.idx ] || bash $ROOT/test_scripts/process_dataset_gpt.sh
$LAUNCH $ROOT/pretrain_gpt.py $MODEL_ARGS $TRAIN_ARGS $TOK_ARGS --data-path $DATA_PREFIX --pipeline-model-parallel-size 2 \
  --num-layers-per-virtual-pipeline-stage 1 --use-dpp ${DPP_POLICY:+--dpp-policy $DPP_POLICY} "$@"
