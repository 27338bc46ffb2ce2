# round 2, call 30 (1 GPU): bench.py after adding the in-run cuBLAS int8 comparator to the roofline block
mkdir -p gpurun_out
(OZ_BENCH_CONFIG4=0 timeout 900 python bench.py --steps 10 --warmup 3) > gpurun_out/r2_bench_ours_d.json 2> gpurun_out/r2_bench_ours_d.err; echo "bench ours rc=$?"; tail -3 gpurun_out/r2_bench_ours_d.err
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_ours_d.json').read().strip().splitlines()[-1])
print(round(d['value'],2), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'roofline', d['roofline'])"
timeout 120 python bench.py 2>&1 | tail -c 600
