# round 2, call 2 (1 GPU): the whole GPU suite after the queue removal + new tests; 4096 default-dispatch check
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q --maxfail=15) > gpurun_out/r2_t2.log 2>&1; echo "pytest gpu rc=$?"; tail -40 gpurun_out/r2_t2.log
timeout 200 python tools/perf_probe.py 4096 9 --iters 8 --shapes 00,p256,00,p192,00 --zgemm --no-extras 2>&1 | tee gpurun_out/r2_perf_4096b.txt
timeout 200 python tools/perf_probe.py 1024 9 --iters 20 --shapes 00,p256,p192,p128 --no-extras 2>&1 | tee gpurun_out/r2_perf_1024.txt
timeout 200 python tools/perf_probe.py 2048 9 --iters 20 --shapes 00,p256,p192,p128 --no-extras 2>&1 | tee gpurun_out/r2_perf_2048.txt
