mkdir -p gpurun_out
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:oz_gemm_pair -c 1 -o gpurun_out/prof_pair192_r1d python tools/perf_probe.py 8192 9 --iters 1) > gpurun_out/ncu4.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu4.log
