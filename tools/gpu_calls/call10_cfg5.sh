# BASELINE config #5: Llama-3 8B, TP8, fp8, MegaScan slow-rank injection + detector
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
COMMON="--gpus 8 --model llama3_8b --layout tp8 --micro-batch 2 --global-batch 16 --steps 4 --warmup 3"
timeout 300 $TR --master-port 29561 bench.py $COMMON --fp8 --detect > gpurun_out/r2_cfg5_fp8_healthy.log 2>&1; echo "fp8 healthy rc=$?"
tail -1 gpurun_out/r2_cfg5_fp8_healthy.log | cut -c1-600
timeout 300 $TR --master-port 29562 bench.py $COMMON --fp8 --slow-rank-inject 5:350000 --detect > gpurun_out/r2_cfg5_fp8_slow5.log 2>&1; echo "fp8 slow rc=$?"
tail -1 gpurun_out/r2_cfg5_fp8_slow5.log | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print({k:d.get(k) for k in ('ms_per_step','trace_overhead_pct','detector','slow_rank_inject')})
except Exception as e: print('parse fail', e)
"
timeout 300 $TR --master-port 29563 bench.py $COMMON > gpurun_out/r2_cfg5_bf16.log 2>&1; echo "bf16 rc=$?"
tail -1 gpurun_out/r2_cfg5_bf16.log | cut -c1-600
grep -h -i "error\|Traceback" -A3 gpurun_out/r2_cfg5_*.log | head -30
