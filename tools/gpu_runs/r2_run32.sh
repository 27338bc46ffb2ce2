# round 2, call 32 (4 GPUs): bench --gpus 4 with the final code
mkdir -p gpurun_out
(timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29632 bench.py --gpus 4 --steps 10 --warmup 3) > gpurun_out/r2_bench_4gpu_c.json 2> gpurun_out/r2_bench_4gpu_c.err; echo "bench N=4 rc=$?"; tail -2 gpurun_out/r2_bench_4gpu_c.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_4gpu_c.json').read().strip().splitlines()[-1])
print('N=4 value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'per_rank', [round(x,2) for x in d['per_rank_ms']], 'bcast', d['bcast_ms'], 'parity', d['parity']['max_ulp'], 'accuracy', d['accuracy']['max_rel_err_vs_cublas_dgemm'], 'e2e', round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],2), [round(x,1) for x in d['e2e']['per_rank_ms']], 'config4', round(d['config4']['value'],1), round(d['config4']['ms_per_step'],2), d['config4']['parity']['max_ulp'])
PY
