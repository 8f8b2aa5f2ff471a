#!/bin/bash
# HF GPT-2 / Llama / Mistral / Mixtral directory -> this framework's checkpoint (reference convert_hf_to_megatron.sh)
source "$(dirname "$0")/_env.sh"
python $ROOT/tools/checkpoint/convert.py --model-type ${MODEL_TYPE:-GPT} --loader ${LOADER:-hf_gpt2} --saver core \
  --load-dir ${HF_DIR:?set HF_DIR} --save-dir ${SAVE_DIR:-$WORK/converted} --target-tensor-parallel-size ${TP:-1} \
  --target-pipeline-parallel-size ${PP:-1} "$@"
