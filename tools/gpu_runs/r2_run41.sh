# round 2, call 41 (1 GPU): evidence of the FINAL commit: GPU suite, smoke, both bench arms (reference first), launch
# list, ncu --set full of the fused kernel
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q --maxfail=10) > gpurun_out/r2_t41.log 2>&1; echo "pytest gpu rc=$?"; grep -E "passed|failed" gpurun_out/r2_t41.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
(timeout 600 python bench.py --impl reference --steps 10 --warmup 3) > gpurun_out/r2_bench_reference_h.json 2> gpurun_out/r2_bench_reference_h.err; echo "bench ref rc=$?"
(timeout 900 python bench.py --steps 10 --warmup 3) > gpurun_out/r2_bench_ours_h.json 2> gpurun_out/r2_bench_ours_h.err; echo "bench ours rc=$?"; tail -3 gpurun_out/r2_bench_ours_h.err
python -c "
import json
for f in ('ours','reference'):
    d=json.loads(open('gpurun_out/r2_bench_%s_h.json'%f).read().strip().splitlines()[-1])
    print(f, round(d['value'],2), 'TFLOP/s', round(d['ms_per_step'],2),'ms; e2e', round(d['e2e']['value'],2), round(d['e2e'].get('ms_per_step'),2), 'roof', (d.get('roofline') or {}).get('frac'), (d.get('roofline') or {}).get('frac_of_cublas_int8'), 'launches', d.get('gpu_launches'), 'config4', (d.get('config4') or {}).get('value'), 'clocks', d['clocks'])"
(OZ_BENCH_CONFIG4=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_bench_launches_h.csv python bench.py --steps 2 --warmup 1) > gpurun_out/r2_bench_under_ncu_h.log 2>&1; echo "ncu launches rc=$?"
(timeout 900 ncu --set full --clock-control none --import-source on -k regex:oz_gemm_pair -c 1 -f -o gpurun_out/r2_prof_pair256_h python tools/perf_probe.py 8192 9 --iters 1 --no-extras) > gpurun_out/r2_ncu_full_h.log 2>&1; echo "ncu full rc=$?"
