# round 2, call 51 (2 GPUs): after removing the split-only panel form: sharded tests + bench --gpus 2
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_sharded.py -m gpu -q -x) > gpurun_out/r2_t51.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2_t51.log
(OZ_BENCH_CONFIG4=0 timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29651 bench.py --gpus 2 --steps 5 --warmup 3) > gpurun_out/r2_bench_2gpu_m.json 2> gpurun_out/r2_bench_2gpu_m.err; echo "bench N=2 rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_2gpu_m.json').read().strip().splitlines()[-1])
print(d['n_gpus'], round(d['value'],2), 'TFLOP/s', round(d['ms_per_step'],2),'ms per_rank', [round(x,2) for x in d['per_rank_ms']], 'e2e', round(d['e2e'].get('ms_per_step',0),2), d['e2e'].get('bit_identical_to_device_path'), 'parity', d['parity']['max_ulp'])"
