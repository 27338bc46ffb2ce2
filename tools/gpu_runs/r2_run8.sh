# round 2, call 8 (8 GPUs): sharded parity test, bench at 8 / 4 / 2 GPUs (weak scaling + config 4 + parity + e2e), config-4 shaped probe
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
(timeout 300 python -m pytest tests/test_gpu_sharded.py -m gpu -q -x) > gpurun_out/r2_t8.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_t8.log
for N in 8 4 2; do
  (timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29550+N)) bench.py --gpus $N --steps 10 --warmup 3) > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err; echo "bench N=$N rc=$?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_${N}gpu.json').read().strip().splitlines()[-1])
print('N=$N value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'per_rank', [round(x,2) for x in d['per_rank_ms']], 'bcast', d['bcast_ms'], 'parity', d['parity']['max_ulp'], 'e2e', round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],2), d['e2e'].get('bit_identical_to_device_path'), 'config4', round(d['config4']['value'],1), round(d['config4']['ms_per_step'],2), d['config4']['parity'])
PY
done
(timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 tools/sharded_probe.py 16384 2048 panels=1,2,3,4 nohost) 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM\|^$" | tee gpurun_out/r2_sharded_probe_8gpu_config4.txt
