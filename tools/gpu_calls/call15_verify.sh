mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_kernels_gpu.py -x -q -k "paged or moe_layer or loss_curve or model_step or smoke" > gpurun_out/r2_pytest_gpu_tail.log 2>&1; echo "pytest tail rc=$?"
tail -6 gpurun_out/r2_pytest_gpu_tail.log | cut -c1-600
