# round 2, call 20 (1 GPU): validation of the tree after the small-problem tile, the k-aware dispatch and the A-first host
# pipeline: GPU suite, smoke, both bench arms (reference first), launch list, ncu --set full of the fused kernel,
# config-3 and config-5 sweeps
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q --maxfail=10) > gpurun_out/r2_t20.log 2>&1; echo "pytest gpu rc=$?"; tail -4 gpurun_out/r2_t20.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
(timeout 600 python bench.py --impl reference --steps 10 --warmup 3) > gpurun_out/r2_bench_reference_b.json 2> gpurun_out/r2_bench_reference_b.err; echo "bench ref rc=$?"
(timeout 900 python bench.py --steps 10 --warmup 3) > gpurun_out/r2_bench_ours_b.json 2> gpurun_out/r2_bench_ours_b.err; echo "bench ours rc=$?"; tail -3 gpurun_out/r2_bench_ours_b.err
python -c "
import json
for f in ('ours','reference'):
    d=json.loads(open('gpurun_out/r2_bench_%s_b.json'%f).read().strip().splitlines()[-1])
    print(f, round(d['value'],2), 'TFLOP/s', round(d['ms_per_step'],2),'ms; e2e', round(d['e2e']['value'],2), d['e2e'].get('ms_per_step'), 'roof', (d.get('roofline') or {}).get('frac'), 'launches', d.get('gpu_launches'), 'clocks', d['clocks'], 'config4', (d.get('config4') or {}).get('value'), (d.get('config4') or {}).get('ms_per_step'), 'cpu', (d.get('cpu_baseline') or {}).get('value'))"
(OZ_BENCH_CONFIG4=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_bench_launches_b.csv python bench.py --steps 2 --warmup 1) > gpurun_out/r2_bench_under_ncu_b.log 2>&1; echo "ncu launches rc=$?"; wc -l gpurun_out/r2_bench_launches_b.csv
(timeout 900 ncu --set full --clock-control none --import-source on -k regex:oz_gemm_pair -c 1 -f -o gpurun_out/r2_prof_pair256_b python tools/perf_probe.py 8192 9 --iters 1 --no-extras) > gpurun_out/r2_ncu_full_b.log 2>&1; echo "ncu full rc=$?"
timeout 900 python tools/sweeps.py split 8192 2>&1 | tee gpurun_out/r2_config3_split_sweep_8192_b.csv | tail -20
timeout 900 python tools/sweeps.py auto 4096 2>&1 | tee gpurun_out/r2_config5_auto_sweep_4096_b.csv | tail -12
