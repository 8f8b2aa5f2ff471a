# BASELINE config #3: Mixtral 8x7B, TP2 x EP4, 8 GPUs -- own arm (arena dispatcher + grouped tcgen05 GEMM) vs the in-tree
# reference algorithm (NCCL all-to-all dispatcher + per-expert GEMMs)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
COMMON="--gpus 8 --model mixtral_8x7b --layout tp2ep4 --micro-batch 1 --global-batch 16 --steps 4 --warmup 3"
timeout 420 $TR --master-port 29551 bench.py $COMMON > gpurun_out/r2_mixtral_own.log 2>&1; echo "own rc=$?"
tail -1 gpurun_out/r2_mixtral_own.log | cut -c1-1200
timeout 420 $TR --master-port 29552 bench.py $COMMON --moe-baseline > gpurun_out/r2_mixtral_base.log 2>&1; echo "base rc=$?"
tail -1 gpurun_out/r2_mixtral_base.log | cut -c1-700
grep -h -i "error\|Traceback" -A3 gpurun_out/r2_mixtral_own.log gpurun_out/r2_mixtral_base.log | head -30
