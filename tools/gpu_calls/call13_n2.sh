# 2 GPUs: bucket reduce-scatter sweep + K14 embedding scatter, synccheck of the flag-protocol kernels, recompute + arena MoE,
# data-parallel 6.7B (ZeRO-1, overlapped grad reduce-scatter: NCCL vs the fused kernel)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 120 $TR --master-port 29581 tools/multi_gpu_checks.py --dp-only --skip-model --skip-moe --out gpurun_out/r2_dp_embed_n2.json > gpurun_out/r2_dp_embed_n2.log 2>&1; echo "dp/embed rc=$?"
grep -E "bucket_rs|embed|EXCEPTION" gpurun_out/r2_dp_embed_n2.log | cut -c1-900
timeout 110 $TR --master-port 29582 --no-python compute-sanitizer --tool synccheck --log-file gpurun_out/r2_synccheck_%p.log python tools/sanitize_comm.py > gpurun_out/r2_synccheck_run.log 2>&1; echo "synccheck rc=$?"
grep -h "OK\|DONE" gpurun_out/r2_synccheck_run.log | head -12; grep -h "ERROR SUMMARY\|Barrier error\|error" gpurun_out/r2_synccheck_[0-9]*.log | sort | uniq -c | head
timeout 120 $TR --master-port 29583 bench.py --gpus 2 --model tiny_moe --layout tp1ep2 --recompute-layers 1 --steps 3 --warmup 3 > gpurun_out/r2_tinymoe_recompute.log 2>&1; echo "tiny_moe recompute rc=$?"
tail -1 gpurun_out/r2_tinymoe_recompute.log | cut -c1-400
timeout 150 $TR --master-port 29584 bench.py --gpus 2 --layout tp1dp2 --micro-batch 1 --steps 3 --warmup 3 > gpurun_out/r2_bench_dp2_nccl.log 2>&1; echo "dp2 nccl rc=$?"
tail -1 gpurun_out/r2_bench_dp2_nccl.log | cut -c1-700
timeout 150 $TR --master-port 29585 bench.py --gpus 2 --layout tp1dp2 --micro-batch 1 --steps 3 --warmup 3 --fused-grad-rs > gpurun_out/r2_bench_dp2_fused.log 2>&1; echo "dp2 fused rc=$?"
tail -1 gpurun_out/r2_bench_dp2_fused.log | cut -c1-700
grep -h "Error\|error" -A2 gpurun_out/r2_bench_dp2_*.log gpurun_out/r2_tinymoe_recompute.log | head -20
