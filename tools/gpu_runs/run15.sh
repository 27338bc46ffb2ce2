mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_kernels.py -q --maxfail=5 -k "split") > gpurun_out/t_split.log 2>&1; echo "split rc=$?"; tail -4 gpurun_out/t_split.log
(timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_complex.py -q --maxfail=5) > gpurun_out/t_gemm.log 2>&1; echo "gemm rc=$?"; tail -4 gpurun_out/t_gemm.log
timeout 200 python tools/perf_probe.py 8192 9 --iters 10 2>&1 | head -6
timeout 200 python tools/perf_probe.py 4096 9 --iters 10 2>&1 | head -6
