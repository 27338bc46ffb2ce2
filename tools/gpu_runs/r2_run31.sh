# round 2, call 31 (1 GPU): bench.py twice on a fresh box (is the first e2e number of a cold box slower?)
mkdir -p gpurun_out
for i in 1 2; do
(OZ_BENCH_CONFIG4=0 timeout 900 python bench.py --steps 10 --warmup 3) > gpurun_out/r2_bench_ours_e$i.json 2> gpurun_out/r2_bench_ours_e$i.err; echo "bench ours rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_ours_e$i.json').read().strip().splitlines()[-1])
print(round(d['value'],2), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), d['e2e'], 'cublas_int8', d['roofline']['cublas_int8']['tops'], d['roofline']['frac_of_cublas_int8'])"
done
OZIMMU_B200_E2E_TRACE=1 timeout 200 python tools/e2e_probe.py 8192 768:768 2>&1 | grep -v trace | tail -4
