#!/bin/bash
# HF Llama / Mixtral -> core (tp / pp / ep layout of your choice) -> train -> back to HF; TRT-LLM checkpoint export.
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/../.." && pwd)"
python $ROOT/tools/checkpoint/convert.py --loader ${LOADER:-hf_llama} --saver core --load-dir $HF_DIR --save-dir $CORE_DIR \
  --target-tensor-parallel-size ${TP:-4} --target-pipeline-parallel-size ${PP:-2} --target-expert-parallel-size ${EP:-1}
python $ROOT/tools/checkpoint/convert.py --loader core --saver ${SAVER:-hf_llama} --load-dir $CORE_DIR --save-dir ${HF_OUT:-$HF_DIR.out}
