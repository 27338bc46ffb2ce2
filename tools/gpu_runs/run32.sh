mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_batched.py tests/test_gpu_kernels.py tests/test_gpu_host_blocks.py -x -q) > gpurun_out/t_batched2.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/t_batched2.log
(timeout 600 python tools/batched_probe.py) > gpurun_out/batched_probe2.log 2>&1; echo "probe rc=$?"; cat gpurun_out/batched_probe2.log
(timeout 300 python tools/e2e_probe.py 8192 768:768 512:512 768:768 512:512 640:640 768:768 512:512) 2>&1 | head -8
