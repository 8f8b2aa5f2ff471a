#!/bin/bash
# Inverse-cloze-task pre-training of the BERT dual encoder, then ORQA retrieval accuracy on NQ-open style files.
source "$(dirname "$0")/../_common.sh"
$LAUNCH $ROOT/pretrain_ict.py \
  --num-layers 12 --hidden-size 768 --num-attention-heads 12 --seq-length 256 --max-position-embeddings 512 \
  --micro-batch-size 32 --global-batch-size 256 --train-iters ${TRAIN_ITERS:-100000} --lr 1e-4 --min-lr 1e-5 \
  --lr-decay-style linear --weight-decay 1e-2 --clip-grad 1.0 --bf16 --biencoder-projection-dim 128 \
  --retriever-score-scaling --retriever-report-topk-accuracies 1 5 20 --query-in-block-prob 0.1 --mock-data \
  --tokenizer-type NullTokenizer --vocab-size ${VOCAB_SIZE:-30592} --log-interval 10 --save $CHECKPOINT_PATH "$@"
# evaluation:  $LAUNCH $ROOT/tasks/main.py --task RETRIEVER-EVAL --evidence-data-path psgs_w100.tsv --qa-data-test nq-test.csv ...
