# round 2, call 21 (2 GPUs): sharded parity tests and bench --gpus 2 with the final code (accuracy block included)
mkdir -p gpurun_out
(timeout 400 python -m pytest tests/test_gpu_sharded.py -m gpu -q -x) > gpurun_out/r2_t21.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_t21.log
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 2 --steps 10 --warmup 3) > gpurun_out/r2_bench_2gpu_c.json 2> gpurun_out/r2_bench_2gpu_c.err; echo "bench N=2 rc=$?"; tail -2 gpurun_out/r2_bench_2gpu_c.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_2gpu_c.json').read().strip().splitlines()[-1])
print('N=2 value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'per_rank', [round(x,2) for x in d['per_rank_ms']], 'bcast', d['bcast_ms'], 'parity', d['parity']['max_ulp'], 'accuracy', d['accuracy'], 'e2e', round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],2), 'config4', round(d['config4']['value'],1), round(d['config4']['ms_per_step'],2))
PY
