mkdir -p gpurun_out
(timeout 100 python -m pytest tests/test_gpu_gemm.py -x -q -k "matches_oracle or headline or cluster_shapes") 2>&1 | tail -2
(timeout 60 python tools/perf_probe.py 8192 9 --iters 10) 2>&1 | head -1
(OZIMMU_B200_TEST_QUEUE=1 timeout 75 python -m pytest tests/test_gpu_queue.py -x -q) > gpurun_out/t_queue.log 2>&1; echo "queue tests rc=$?"; tail -15 gpurun_out/t_queue.log
(OZIMMU_B200_E2E_QUEUE=1 timeout 50 python tools/e2e_probe.py 8192 768:768 1024:1024) 2>&1 | head -3
