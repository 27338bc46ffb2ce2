# round 2, call 29 (1 GPU): final validation of the committed tree: GPU suite, smoke, both bench arms (reference first)
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q --maxfail=10) > gpurun_out/r2_t29.log 2>&1; echo "pytest gpu rc=$?"; grep -E "passed|failed" gpurun_out/r2_t29.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
(timeout 600 python bench.py --impl reference --steps 10 --warmup 3) > gpurun_out/r2_bench_reference_c.json 2> gpurun_out/r2_bench_reference_c.err; echo "bench ref rc=$?"
(timeout 900 python bench.py --steps 10 --warmup 3) > gpurun_out/r2_bench_ours_c.json 2> gpurun_out/r2_bench_ours_c.err; echo "bench ours rc=$?"; tail -3 gpurun_out/r2_bench_ours_c.err
python -c "
import json
for f in ('ours','reference'):
    d=json.loads(open('gpurun_out/r2_bench_%s_c.json'%f).read().strip().splitlines()[-1])
    print(f, round(d['value'],2), 'TFLOP/s', round(d['ms_per_step'],2),'ms; e2e', round(d['e2e']['value'],2), round(d['e2e'].get('ms_per_step'),2), 'roof', (d.get('roofline') or {}).get('frac'), 'launches', d.get('gpu_launches'), 'accuracy', d.get('accuracy'), 'config4', (d.get('config4') or {}).get('value'))"
