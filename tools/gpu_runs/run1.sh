mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt
(timeout 300 python tests/golden/make_golden.py gpurun_out/golden) > gpurun_out/golden.log 2>&1; echo "golden rc=$?"
(timeout 600 python -m pytest tests/test_gpu_kernels.py -q --maxfail=8 -x -k "not pair_product") > gpurun_out/t_kernels_split.log 2>&1; echo "split rc=$?"; tail -5 gpurun_out/t_kernels_split.log
(timeout 300 python -m pytest tests/test_gpu_kernels.py -q --maxfail=4 -k "pair_product") > gpurun_out/t_kernels_pair.log 2>&1; echo "pair rc=$?"; tail -15 gpurun_out/t_kernels_pair.log
(timeout 600 python -m pytest tests/test_gpu_gemm.py -q --maxfail=6) > gpurun_out/t_gemm.log 2>&1; echo "gemm rc=$?"; tail -15 gpurun_out/t_gemm.log
(timeout 300 python tools/perf_probe.py 4096 9 --ref --shapes 11,21,12,22) > gpurun_out/perf4096.log 2>&1; echo "perf rc=$?"; cat gpurun_out/perf4096.log
