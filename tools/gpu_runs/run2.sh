mkdir -p gpurun_out
(timeout 300 python tests/golden/make_golden.py gpurun_out/golden) > gpurun_out/golden.log 2>&1; echo "golden rc=$?"; tail -3 gpurun_out/golden.log
(timeout 900 python -m pytest tests/test_gpu_gemm.py -q --maxfail=6) > gpurun_out/t_gemm.log 2>&1; echo "gemm rc=$?"; tail -15 gpurun_out/t_gemm.log
(timeout 600 python -m pytest tests/test_gpu_auto_and_dropin.py -q --maxfail=6) > gpurun_out/t_auto.log 2>&1; echo "auto rc=$?"; tail -25 gpurun_out/t_auto.log
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:oz_gemm_fused -c 1 -o gpurun_out/prof_fused_r1a python tools/perf_probe.py 4096 9 --iters 1) > gpurun_out/ncu1.log 2>&1; echo "ncu rc=$?"; tail -5 gpurun_out/ncu1.log
(timeout 300 python tools/perf_probe.py 8192 9 --ref --shapes 11,21) > gpurun_out/perf8192.log 2>&1; echo "perf rc=$?"; cat gpurun_out/perf8192.log
