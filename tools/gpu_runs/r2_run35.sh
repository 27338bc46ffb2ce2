# round 2, call 35 (1 GPU): streamlined producer loop (unpaced fast path, running pointers, one B piece at compile time)
# and plain wait on the relayed barrier: full GPU suite, then sizes 1024 .. 8192 and the 128 x 128 kernel time
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q --maxfail=10) > gpurun_out/r2_t35.log 2>&1; echo "pytest gpu rc=$?"; grep -E "passed|failed" gpurun_out/r2_t35.log
for n in 1024 1536 2048 4096 8192; do
  timeout 300 python tools/perf_probe.py $n 9 --iters 20 --shapes 00,00 --no-extras 2>&1 | tee -a gpurun_out/r2_perf_streamlined.txt
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_1024_c.csv python tools/perf_probe.py 1024 9 --iters 2 --shapes h128,p128 --no-extras > /dev/null 2>&1
python - <<PY
import csv
rows = [r for r in csv.reader(open('gpurun_out/r2_launches_1024_c.csv')) if len(r) > 5 and r[0].isdigit()]
for r in rows[-24:]:
    if 'pair' in r[4]: print(r[4][:90], r[-1], r[-2])
PY
