#!/bin/bash
# start the REST text-generation server on a checkpoint, query it once, stop it (MegaScope: add --enable-ws-server)
source "$(dirname "$0")/_env.sh"
SERVER_PORT=${SERVER_PORT:-5000}
NPROC=1
python $ROOT/tools/run_text_generation_server.py $MODEL_ARGS $TOK_ARGS --distributed-backend $BACKEND ${LOAD:+--load $LOAD} --port $SERVER_PORT "$@" &
PID=$!
trap "kill $PID 2>/dev/null" EXIT
for i in $(seq 1 60); do
  sleep 2
  python $ROOT/tools/text_generation_cli.py 127.0.0.1:$SERVER_PORT --prompt "Northwind Holdings reports" --tokens-to-generate 8 && exit 0
done
echo "server did not answer" >&2; exit 1
