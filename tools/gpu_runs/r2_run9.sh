# round 2, call 9 (8 GPUs): bench at 8 / 4 / 2 GPUs after "owner uploads B first" + NUMA binding (e2e at N > 1), sharded test
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_sharded.py -m gpu -q -x) > gpurun_out/r2_t9.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_t9.log
for N in 8 4 2; do
  (OZ_BENCH_CONFIG4=$([ $N = 8 ] && echo 1 || echo 0) timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29570+N)) bench.py --gpus $N --steps 10 --warmup 3) > gpurun_out/r2_bench_${N}gpu_b.json 2> gpurun_out/r2_bench_${N}gpu_b.err; echo "bench N=$N rc=$?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_${N}gpu_b.json').read().strip().splitlines()[-1])
print('N=$N value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'per_rank', [round(x,2) for x in d['per_rank_ms']], 'bcast', d['bcast_ms'], 'parity', d['parity']['max_ulp'], 'e2e', round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],2), [round(x,1) for x in d['e2e']['per_rank_ms']], d['e2e'].get('bit_identical_to_device_path'))
PY
done
