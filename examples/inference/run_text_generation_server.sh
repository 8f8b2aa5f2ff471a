#!/bin/bash
# REST (PUT /api, POST /v1/completions) + optional MegaScope WebSocket server.
source "$(dirname "$0")/../_common.sh"
GPUS_PER_NODE=${GPUS_PER_NODE:-1}
python -m torch.distributed.run --nproc-per-node $GPUS_PER_NODE --master-addr $MASTER_ADDR --master-port $MASTER_PORT \
  $ROOT/tools/run_text_generation_server.py --tensor-model-parallel-size $GPUS_PER_NODE \
  --num-layers 32 --hidden-size 4096 --num-attention-heads 32 --seq-length 4096 --max-position-embeddings 4096 \
  --tokenizer-type ${TOKENIZER:-GPT2BPETokenizer} --vocab-file $VOCAB_FILE --merge-file $MERGE_FILE \
  --bf16 --micro-batch-size 1 --load $CHECKPOINT_PATH --port ${PORT:-5000} ${ENABLE_WS:+--enable-ws-server} "$@"
