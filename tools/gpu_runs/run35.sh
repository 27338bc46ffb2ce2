mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29560 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err; echo "bench 8 rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/bench_8gpu.json').read().strip().splitlines()[-1])
print('N=8', round(d['value'],2), 'TFLOP/s', round(d['ms_per_step'],3),'ms; e2e', round(d['e2e']['value'],2))"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 tools/ubench/nccl_bcast.py 2>&1 | grep "^world" | tee gpurun_out/nccl_bcast_8gpu.log
