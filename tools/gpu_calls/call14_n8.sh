# 8 GPUs, one box: BASELINE configs #5 (Llama-3 8B TP8 fp8 + slow-rank detection), #4 (GPT-3 1.3B TP1 PP4: static interleaved
# vs MegaDPP vs MegaFBD interleaved) and #3 (Mixtral 8x7B TP2 EP4, own arm; partial activation recompute to fit 180 GB)
mkdir -p gpurun_out
export PYTORCH_CUDA_ALLOC_CONF=expandable_segments:True
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
C5="--gpus 8 --model llama3_8b --layout tp8 --micro-batch 2 --global-batch 16 --steps 3 --warmup 3"
timeout 150 $TR --master-port 29562 bench.py $C5 --fp8 --slow-rank-inject 5:350000 --detect > gpurun_out/r2_cfg5_fp8_slow5.log 2>&1; echo "cfg5 fp8 slow rc=$?"
tail -1 gpurun_out/r2_cfg5_fp8_slow5.log | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print({k:d.get(k) for k in ('ms_per_step','trace_overhead_pct','detector','slow_rank_inject')})
except Exception as e: print('parse fail', e)
"
C4="--gpus 8 --model gpt3_1.3b --micro-batch 4 --steps 4 --warmup 3"
timeout 120 $TR --master-port 29571 bench.py $C4 --layout tp1pp4dp2 --vpp 2 > gpurun_out/r2_cfg4_static.log 2>&1; echo "cfg4 static rc=$?"
tail -1 gpurun_out/r2_cfg4_static.log | cut -c1-300
timeout 120 $TR --master-port 29573 bench.py $C4 --layout tp1pp4 --vpp 2 --fbd > gpurun_out/r2_cfg4_fbd.log 2>&1; echo "cfg4 fbd rc=$?"
tail -1 gpurun_out/r2_cfg4_fbd.log | cut -c1-300
timeout 240 $TR --master-port 29551 bench.py --gpus 8 --model mixtral_8x7b --layout tp2ep4 --micro-batch 1 --global-batch 16 --steps 3 --warmup 3 --recompute-layers 16 > gpurun_out/r2_mixtral_own.log 2>&1; echo "mixtral own rc=$?"
tail -1 gpurun_out/r2_mixtral_own.log | cut -c1-900
grep -h "Error" gpurun_out/r2_cfg5_*.log gpurun_out/r2_cfg4_*.log gpurun_out/r2_mixtral_own.log | sort | uniq -c | head -10
