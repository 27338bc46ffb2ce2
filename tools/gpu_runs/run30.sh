nvidia-smi -L | wc -l
(timeout 600 python -m pytest tests/test_gpu_sharded.py -q) 2>&1 | tail -2
for pw in 2048 1024 4096; do echo -n "E2E_PANEL=$pw: "; OZIMMU_B200_E2E_PANEL=$pw timeout 200 python tools/e2e_probe.py 8192 2>&1 | head -3 | tr '\n' ' '; echo; done
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3) 2>/dev/null | tail -1 > gpurun_out/bench_2gpu.json; python -c "
import json
d=json.loads(open('gpurun_out/bench_2gpu.json').read())
print(d['n_gpus'], round(d['value'],2), 'TFLOP/s', round(d['ms_per_step'],2),'ms; e2e', round(d['e2e']['value'],2))"
