# round 2, call 45 (8 GPUs): bench --gpus 8 with the final bench.py (clock sampler before the barrier)
mkdir -p gpurun_out
(timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29645 bench.py --gpus 8 --steps 10 --warmup 3) > gpurun_out/r2_bench_8gpu_k.json 2> gpurun_out/r2_bench_8gpu_k.err; echo "bench N=8 rc=$?"; tail -2 gpurun_out/r2_bench_8gpu_k.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_8gpu_k.json').read().strip().splitlines()[-1])
print('N=8 value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'per_rank', [round(x,2) for x in d['per_rank_ms']], 'bcast', d['bcast_ms'], 'parity', d['parity']['max_ulp'], 'e2e', round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],2), 'config4', round(d['config4']['value'],1), round(d['config4']['ms_per_step'],2), d['config4']['parity']['max_ulp'], 'roof', round(d['roofline']['frac'],3), d['roofline']['kernel_ms'])
PY
