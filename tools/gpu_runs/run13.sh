mkdir -p gpurun_out
nvidia-smi -L | wc -l
(timeout 900 python -m pytest tests/test_gpu_auto_and_dropin.py tests/test_gpu_sharded.py -q --maxfail=5) > gpurun_out/t_new.log 2>&1; echo "new tests rc=$?"; tail -12 gpurun_out/t_new.log
for n in 2 4 8; do
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 5 --warmup 3) > gpurun_out/bench_${n}gpu.json 2> gpurun_out/bench_${n}gpu.err; echo "bench$n rc=$?"; python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_${n}gpu.json').read().strip().splitlines()[-1])
print(d['n_gpus'], round(d['value'],2), 'TFLOP/s', round(d['ms_per_step'],2),'ms; e2e', round(d['e2e']['value'],2), 'clocks', d['clocks'])
"; tail -2 gpurun_out/bench_${n}gpu.err | cut -c1-300
done
(timeout 600 python bench.py --steps 10 --warmup 3) > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "bench ours rc=$?"; python -c "
import json
d=json.loads(open('gpurun_out/bench_ours.json').read().strip().splitlines()[-1])
print(d['n_gpus'], round(d['value'],2), 'TFLOP/s', round(d['ms_per_step'],2),'ms; e2e', round(d['e2e']['value'],2), 'roof', d['roofline']['frac'], 'clocks', d['clocks'])"
