# round 2, call 7 (1 GPU): both bench arms (N=1, with config 4 and parity blocks), launch list, ncu --set full of the
# fused kernel, split kernels after the band / cache-hint change, config-5 sweep
mkdir -p gpurun_out
(timeout 600 python bench.py --impl reference --steps 10 --warmup 3) > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; echo "bench ref rc=$?"
(timeout 900 python bench.py --steps 10 --warmup 3) > gpurun_out/r2_bench_ours.json 2> gpurun_out/r2_bench_ours.err; echo "bench ours rc=$?"; tail -3 gpurun_out/r2_bench_ours.err
python -c "
import json
for f in ('ours','reference'):
    d=json.loads(open('gpurun_out/r2_bench_%s.json'%f).read().strip().splitlines()[-1])
    print(f, round(d['value'],2), 'TFLOP/s', round(d['ms_per_step'],2),'ms; e2e', round(d['e2e']['value'],2), d['e2e'].get('ms_per_step'), 'roof', (d.get('roofline') or {}).get('frac'), 'launches', d.get('gpu_launches'), 'clocks', d['clocks'], 'config4', (d.get('config4') or {}).get('value'), (d.get('config4') or {}).get('ms_per_step'), 'cpu', (d.get('cpu_baseline') or {}).get('value'))"
(OZ_BENCH_CONFIG4=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 2 --warmup 1) > gpurun_out/r2_bench_under_ncu.log 2>&1; echo "ncu launches rc=$?"; wc -l gpurun_out/r2_bench_launches.csv
(timeout 900 ncu --set full --clock-control none --import-source on -k regex:oz_gemm_pair -c 1 -o gpurun_out/r2_prof_pair256 python tools/perf_probe.py 8192 9 --iters 1 --no-extras) > gpurun_out/r2_ncu_full.log 2>&1; echo "ncu full rc=$?"; ls -la gpurun_out/*.ncu-rep
(timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:split -c 4 --csv --log-file gpurun_out/r2_split_kernels_ncu.csv python tools/perf_probe.py 8192 9 --iters 1 --no-extras) > gpurun_out/r2_ncu_split.log 2>&1; python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/r2_split_kernels_ncu.csv")) if len(r) > 10 and r[0].isdigit()]
for r in rows:
    print(r[0], r[4].split("(")[0][-40:], r[-3], r[-1])
PY
timeout 900 python tools/sweeps.py auto 4096 2>&1 | tee gpurun_out/r2_config5_auto_sweep_4096.csv | tail -30
