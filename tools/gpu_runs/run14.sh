mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_gemm.py -q -k "long_k") > gpurun_out/t_longk.log 2>&1; echo "longk rc=$?"; tail -3 gpurun_out/t_longk.log
(timeout 900 python tools/sweeps.py split 8192) > gpurun_out/split_sweep.csv 2> gpurun_out/split_sweep.err; echo "split sweep rc=$?"; cat gpurun_out/split_sweep.csv; tail -2 gpurun_out/split_sweep.err
(timeout 900 python tools/sweeps.py auto 4096) > gpurun_out/auto_sweep.csv 2> gpurun_out/auto_sweep.err; echo "auto sweep rc=$?"; cut -d, -f1-7 gpurun_out/auto_sweep.csv; tail -2 gpurun_out/auto_sweep.err
