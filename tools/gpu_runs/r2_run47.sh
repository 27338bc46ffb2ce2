# round 2, call 47 (1 GPU): strided split with 16-column row-max CTAs for small matrices: split tests, 1024 / 1536 / 2048
# timings and the split kernel's own time
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_gemm.py tests/test_gpu_complex.py -m gpu -q --maxfail=10) > gpurun_out/r2_t47.log 2>&1; echo "pytest gpu rc=$?"; grep -E "passed|failed" gpurun_out/r2_t47.log
for n in 1024 1536 2048; do
  timeout 300 python tools/perf_probe.py $n 9 --iters 20 --shapes 00,00 --no-extras --graph 2>&1 | tee -a gpurun_out/r2_perf_split_small.txt
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_1024_e.csv python tools/perf_probe.py 1024 9 --iters 2 --shapes 00 --no-extras > /dev/null 2>&1
python - <<PY
import csv
rows = [r for r in csv.reader(open('gpurun_out/r2_launches_1024_e.csv')) if len(r) > 5 and r[0].isdigit()]
for r in rows[-6:]:
    print(r[4][:70], r[-1])
PY
