# round 2, call 44 (2 GPUs): bench --gpus 2 after starting the clock sampler before the barrier (twice)
mkdir -p gpurun_out
for i in 1 2; do
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2966$i bench.py --gpus 2 --steps 10 --warmup 3) > gpurun_out/r2_bench_2gpu_j$i.json 2> gpurun_out/r2_bench_2gpu_j$i.err; echo "bench N=2 rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_2gpu_j$i.json').read().strip().splitlines()[-1])
print(d['n_gpus'], round(d['value'],2), 'TFLOP/s', round(d['ms_per_step'],2),'ms per_rank', [round(x,2) for x in d['per_rank_ms']], 'e2e', round(d['e2e'].get('ms_per_step',0),2), 'parity', d['parity']['max_ulp'], 'config4', round(d['config4']['value'],1), 'clocks', d['clocks']['sm_mhz'], d['clocks']['samples'])"
done
