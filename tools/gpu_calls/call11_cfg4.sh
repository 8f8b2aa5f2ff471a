# BASELINE config #4: GPT-3 1.3B, TP1 PP4 DP2 (8 GPUs), interleaved 1F1B static vs MegaDPP vs MegaFBD
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
COMMON="--gpus 8 --model gpt3_1.3b --micro-batch 4 --steps 6 --warmup 3"
timeout 240 $TR --master-port 29571 bench.py $COMMON --layout tp1pp4dp2 --vpp 2 > gpurun_out/r2_cfg4_static.log 2>&1; echo "static rc=$?"
tail -1 gpurun_out/r2_cfg4_static.log | cut -c1-500
timeout 240 $TR --master-port 29572 bench.py $COMMON --layout tp1pp4dp2 --vpp 2 --dpp auto > gpurun_out/r2_cfg4_dpp.log 2>&1; echo "dpp rc=$?"
tail -1 gpurun_out/r2_cfg4_dpp.log | cut -c1-500
timeout 240 $TR --master-port 29573 bench.py $COMMON --layout tp1pp4 --vpp 2 --fbd > gpurun_out/r2_cfg4_fbd.log 2>&1; echo "fbd rc=$?"
tail -1 gpurun_out/r2_cfg4_fbd.log | cut -c1-500
timeout 240 $TR --master-port 29574 bench.py $COMMON --layout tp1pp4dp2 --vpp 2 --dpp auto --slow-rank-inject 2:300000 > gpurun_out/r2_cfg4_dpp_slow.log 2>&1; echo "dpp slow rc=$?"
tail -1 gpurun_out/r2_cfg4_dpp_slow.log | cut -c1-300
timeout 240 $TR --master-port 29575 bench.py $COMMON --layout tp1pp4dp2 --vpp 2 --slow-rank-inject 2:300000 > gpurun_out/r2_cfg4_static_slow.log 2>&1; echo "static slow rc=$?"
tail -1 gpurun_out/r2_cfg4_static_slow.log | cut -c1-300
grep -h -i "error\|Traceback" -A3 gpurun_out/r2_cfg4_*.log | head -30
