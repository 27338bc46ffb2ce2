mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_kernels.py -q --maxfail=4 -k "pair_product") > gpurun_out/t_kernels_pair.log 2>&1; echo "pair rc=$?"; tail -3 gpurun_out/t_kernels_pair.log
(timeout 900 python -m pytest tests/test_gpu_gemm.py -q --maxfail=6) > gpurun_out/t_gemm.log 2>&1; echo "gemm rc=$?"; tail -3 gpurun_out/t_gemm.log
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 100 > gpurun_out/clocks.csv &
SMI=$!
(timeout 300 python tools/perf_probe.py 4096 9 --shapes p192,p128) > gpurun_out/perf4096.log 2>&1; echo "perf rc=$?"; head -3 gpurun_out/perf4096.log
(timeout 300 python tools/perf_probe.py 8192 9 --shapes p192,p128 --ref --iters 10) > gpurun_out/perf8192.log 2>&1; echo "perf rc=$?"; cat gpurun_out/perf8192.log
kill $SMI
sort gpurun_out/clocks.csv | uniq -c | sort -rn | head -12
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:oz_gemm_pair -c 1 -o gpurun_out/prof_pair192_r1c python tools/perf_probe.py 8192 9 --iters 1) > gpurun_out/ncu3.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu3.log
