mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_kernels.py -q --maxfail=3 -x -k "split") > gpurun_out/t_split.log 2>&1; echo "split rc=$?"; tail -6 gpurun_out/t_split.log
(timeout 200 python -m pytest tests/test_gpu_kernels.py -q --maxfail=3 -x -k "pair_product") > gpurun_out/t_kernels_pair.log 2>&1; echo "pair rc=$?"; tail -12 gpurun_out/t_kernels_pair.log
(timeout 400 python -m pytest tests/test_gpu_gemm.py -q --maxfail=4) > gpurun_out/t_gemm.log 2>&1; echo "gemm rc=$?"; tail -8 gpurun_out/t_gemm.log
(timeout 120 python tools/perf_probe.py 8192 9 --iters 10 --shapes 00,p128) 2>&1 | head -6
(timeout 120 python tools/perf_probe.py 4096 9 --iters 10) 2>&1 | head -2
(timeout 120 python tools/perf_probe.py 1024 9 --iters 30 --shapes 00,p256) 2>&1 | head -3
