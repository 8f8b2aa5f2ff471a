#!/bin/bash
# Retro: build the chunk database + neighbour ids from a .bin/.idx corpus, then pre-train the retrieval-augmented decoder.
source "$(dirname "$0")/../_common.sh"
if [ -n "$DATA_PATH" ]; then
  python $ROOT/tools/retro/build_db.py --data-prefix $DATA_PATH --chunk-length 64 --num-neighbors 2 --retrieved-length 128 --out $ROOT/retro_db
fi
$LAUNCH $ROOT/pretrain_retro.py \
  --tensor-model-parallel-size ${TP:-1} --num-layers 12 --hidden-size 768 --num-attention-heads 12 --seq-length 2048 \
  --max-position-embeddings 2048 --retro-chunk-length 64 --retro-num-neighbors 2 --retro-retrieved-length 128 \
  --retro-encoder-layers 2 --micro-batch-size 4 --global-batch-size 256 --train-iters ${TRAIN_ITERS:-100000} --lr 6e-4 \
  --min-lr 6e-5 --lr-decay-style cosine --weight-decay 0.1 --clip-grad 1.0 --bf16 --mock-data \
  --tokenizer-type NullTokenizer --vocab-size ${VOCAB_SIZE:-50257} --log-interval 10 --save $CHECKPOINT_PATH "$@"
