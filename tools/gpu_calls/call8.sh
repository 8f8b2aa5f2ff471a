mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29541 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/r2_bench_n2c.log 2>&1; echo "n2 rc=$?"
tail -1 gpurun_out/r2_bench_n2c.log | cut -c1-900
timeout 400 $TR --master-port 29542 bench.py --gpus 2 --model llama3_8b --layout tp2 --layers 8 --global-batch 8 --micro-batch 2 --steps 4 --warmup 3 --fp8 --slow-rank-inject 1:400000 --detect > gpurun_out/r2_detect_n2.log 2>&1; echo "detect rc=$?"
tail -1 gpurun_out/r2_detect_n2.log | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print({k:d.get(k) for k in ('ms_per_step','trace_overhead_pct','detector','slow_rank_inject')})
except Exception as e: print('parse fail', e)
"
tail -5 gpurun_out/r2_detect_n2.log | cut -c1-600
