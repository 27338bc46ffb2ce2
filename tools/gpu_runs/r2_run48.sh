# round 2, call 48 (1 GPU): last full GPU suite + smoke + default bench of the final tree
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q --maxfail=10) > gpurun_out/r2_t48.log 2>&1; echo "pytest gpu rc=$?"; grep -E "passed|failed" gpurun_out/r2_t48.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
(timeout 900 python bench.py) > gpurun_out/r2_bench_ours_l.json 2> gpurun_out/r2_bench_ours_l.err; echo "bench ours rc=$?"; tail -3 gpurun_out/r2_bench_ours_l.err
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_ours_l.json').read().strip().splitlines()[-1])
r=d['roofline']
print(round(d['value'],2), 'TFLOP/s', round(d['ms_per_step'],2),'ms; e2e', round(d['e2e']['value'],2), round(d['e2e'].get('ms_per_step'),2), 'roof', round(r['frac'],3), round(r['frac_of_cublas_int8'],3), r['kernel_ms'], 'launches', d.get('gpu_launches'), 'config4', round(d['config4']['value'],1), 'acc', d['accuracy']['max_rel_err_vs_cublas_dgemm'])"
