# round 2, call 40 (2 GPUs): bench.py after the clock-ramp change: N=1 both arms, N=2 under torchrun
mkdir -p gpurun_out
(timeout 600 python bench.py --impl reference --steps 10 --warmup 3) > gpurun_out/r2_bench_reference_g.json 2> gpurun_out/r2_bench_reference_g.err; echo "bench ref rc=$?"
(timeout 900 python bench.py --steps 10 --warmup 3) > gpurun_out/r2_bench_ours_g.json 2> gpurun_out/r2_bench_ours_g.err; echo "bench ours rc=$?"; tail -3 gpurun_out/r2_bench_ours_g.err
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29640 bench.py --gpus 2 --steps 10 --warmup 3) > gpurun_out/r2_bench_2gpu_g.json 2> gpurun_out/r2_bench_2gpu_g.err; echo "bench N=2 rc=$?"; tail -2 gpurun_out/r2_bench_2gpu_g.err
(timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 bench.py --impl reference --gpus 2 --steps 3 --warmup 1) > gpurun_out/r2_bench_ref_2gpu_g.json 2> gpurun_out/r2_bench_ref_2gpu_g.err; echo "bench ref N=2 rc=$?"
python -c "
import json
for f in ('ours_g','reference_g','2gpu_g','ref_2gpu_g'):
    d=json.loads(open('gpurun_out/r2_bench_%s.json'%f).read().strip().splitlines()[-1])
    print(f, d.get('impl'), d['n_gpus'], round(d['value'],2), 'TFLOP/s', round(d['ms_per_step'],2),'ms; e2e', round(d['e2e']['value'],2), round(d['e2e'].get('ms_per_step',0),2), 'launches', d.get('gpu_launches'), 'parity', (d.get('parity') or {}).get('max_ulp'))"
