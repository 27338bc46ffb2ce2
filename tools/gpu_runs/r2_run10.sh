# round 2, call 10 (1 GPU): TMEM layout of M=128 CTA-pair MMAs, GPU suite, where the time of small problems goes
# (ncu launch list + CUDA-graph replay), rasterisation band sweep at 8192^3
mkdir -p gpurun_out
timeout 60 tools/ubench/umma_m128_probe > gpurun_out/r2_umma_m128_probe.txt 2>&1; echo "probe rc=$?"; head -60 gpurun_out/r2_umma_m128_probe.txt
(time timeout 1200 python -m pytest tests -m gpu -q --maxfail=10) > gpurun_out/r2_t10.log 2>&1; echo "pytest gpu rc=$?"; tail -5 gpurun_out/r2_t10.log
for n in 1024 2048; do
  timeout 200 python tools/perf_probe.py $n 9 --iters 20 --shapes 00 --no-extras --graph 2>&1 | tee gpurun_out/r2_perf_${n}_graph.txt
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_${n}.csv python tools/perf_probe.py $n 9 --iters 2 --shapes 00 --no-extras > /dev/null 2>&1
  python - <<PY
import csv
rows = [r for r in csv.reader(open('gpurun_out/r2_launches_${n}.csv')) if len(r) > 5 and r[0].isdigit()]
for r in rows[-12:]:
    print(r[4][:70], r[-1], r[-2])
PY
done
for g in 8 4 6 12 16 32 8; do
  OZIMMU_B200_GROUP_M=$g timeout 200 python tools/perf_probe.py 8192 9 --iters 8 --shapes 00 --no-extras 2>&1 | sed "s/^/group_m=$g /" | tee -a gpurun_out/r2_sweep_group_m.txt
done
