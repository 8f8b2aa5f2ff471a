#!/bin/bash
# PTQ of a trained GPT checkpoint (fp8 | int8_sq | int4_awq | w4a8_awq | nvfp4): calibrate, fold, save + real-weight export.
source "$(dirname "$0")/../_common.sh"
GPUS_PER_NODE=${GPUS_PER_NODE:-1}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $GPUS_PER_NODE --master-addr $MASTER_ADDR --master-port $MASTER_PORT \
  $ROOT/examples/post_training/quantize.py \
  --tensor-model-parallel-size $GPUS_PER_NODE --num-layers 32 --hidden-size 4096 --num-attention-heads 32 --seq-length 2048 \
  --max-position-embeddings 2048 --bf16 --tokenizer-type NullTokenizer --vocab-size ${VOCAB_SIZE:-50257} \
  --load $CHECKPOINT_PATH --save ${CHECKPOINT_PATH}_${QUANT_CFG:-fp8} --export-quant-cfg ${QUANT_CFG:-fp8} --calib-size 128 \
  --calib-batch-size 4 --export-real-weights "$@"
