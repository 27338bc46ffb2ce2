(timeout 200 python -m pytest tests/test_gpu_kernels.py -q --maxfail=3 -x -k "pair_product") 2>&1 | tail -3
(timeout 300 python -m pytest tests/test_gpu_gemm.py -q --maxfail=4 -k "oracle or 8192 or cluster") 2>&1 | tail -3
(timeout 120 python tools/perf_probe.py 8192 9 --iters 10 --shapes 00,p128) 2>&1 | head -6
(timeout 120 python tools/perf_probe.py 4096 9 --iters 10) 2>&1 | head -1
(timeout 120 python tools/perf_probe.py 1024 9 --iters 30 --shapes 00,p256) 2>&1 | head -2
