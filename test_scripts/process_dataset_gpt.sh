#!/bin/bash
# jsonl corpus -> .bin/.idx with the GPT-2 BPE shipped in datasets_gpt/ (reference test_scripts/process_dataset_gpt.sh)
source "$(dirname "$0")/_env.sh"
python $ROOT/tools/preprocess_data.py --input ${INPUT:-$ROOT/datasets_gpt/dataset.json} --output-prefix $WORK/gpt \
  --append-eod --workers ${WORKERS:-2} $TOK_ARGS
ls -l $WORK/gpt_text_document.bin $WORK/gpt_text_document.idx
