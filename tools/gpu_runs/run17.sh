mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/t_all_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -4 gpurun_out/t_all_gpu.log
(timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()") > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -6 gpurun_out/sanitizer_memcheck.log
(timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_kernels.py -q -k "split_matches_oracle and 130-257" ) > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -6 gpurun_out/sanitizer_racecheck.log
(timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"split_|rowmax" -c 12 python tools/perf_probe.py 8192 9 --iters 1) 2>&1 | grep -E "split_|rowmax|gpu__time|dram__" | head -40 | tee gpurun_out/split_ncu.txt
