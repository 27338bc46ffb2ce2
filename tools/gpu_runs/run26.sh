mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_host_blocks.py tests/test_gpu_sharded.py -x -q) > gpurun_out/t_streamed.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/t_streamed.log
for pipe in 0 1; do
OZIMMU_B200_BENCH_PIPELINE=$pipe timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu_pipe$pipe.json 2> gpurun_out/bench_2gpu_pipe$pipe.err; echo "bench pipe=$pipe rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/bench_2gpu_pipe$pipe.json').read().strip().splitlines()[-1])
print('pipe=$pipe', round(d['value'],2), 'TFLOP/s', round(d['ms_per_step'],3),'ms; e2e', round(d['e2e']['value'],2))"
done
