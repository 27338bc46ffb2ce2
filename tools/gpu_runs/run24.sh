mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_batched.py tests/test_gpu_host_blocks.py tests/test_gpu_kernels.py tests/test_gpu_gemm.py -x -q) > gpurun_out/t_batched.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/t_batched.log
(timeout 600 python tools/batched_probe.py) > gpurun_out/batched_probe.log 2>&1; echo "probe rc=$?"; cat gpurun_out/batched_probe.log
(timeout 300 python tools/e2e_probe.py 8192 768:768 1024:1024) 2>&1 | tail -4
