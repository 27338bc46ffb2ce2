# round 2, call 39 (1 GPU): plain wait on the TMEM-buffer barrier in the MMA warp: product-path tests, small sizes
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_kernels.py tests/test_gpu_complex.py tests/test_gpu_batched.py -m gpu -q --maxfail=10) > gpurun_out/r2_t39.log 2>&1; echo "pytest gpu rc=$?"; grep -E "passed|failed" gpurun_out/r2_t39.log
for n in 1024 1536 2048 8192; do
  timeout 300 python tools/perf_probe.py $n 9 --iters 20 --shapes 00,00 --no-extras 2>&1 | tee -a gpurun_out/r2_perf_plainwait.txt
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_1024_d.csv python tools/perf_probe.py 1024 9 --iters 2 --shapes h128 --no-extras > /dev/null 2>&1
python - <<PY
import csv
rows = [r for r in csv.reader(open('gpurun_out/r2_launches_1024_d.csv')) if len(r) > 5 and r[0].isdigit()]
for r in rows[-12:]:
    print(r[4][:70], r[-1])
PY
