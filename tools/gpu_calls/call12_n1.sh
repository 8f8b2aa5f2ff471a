# 1 GPU: the GPU test tier, paged-attention perf, N=1 bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2_pytest_gpu.log
timeout 200 python tools/gpu_checks.py --only paged_attention_perf --out gpurun_out/r2_paged_perf.json > gpurun_out/r2_paged_perf.log 2>&1; echo "paged perf rc=$?"
tail -2 gpurun_out/r2_paged_perf.log | cut -c1-1500
timeout 300 python bench.py --gpus 1 --steps 8 --warmup 3 > gpurun_out/r2_bench_n1b.log 2>&1; echo "bench rc=$?"
tail -1 gpurun_out/r2_bench_n1b.log | cut -c1-1200
