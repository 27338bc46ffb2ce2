# round 2, call 42 (2 GPUs): bench.py after moving the roofline measurement in front of the host-operand loop: N=1, N=2
mkdir -p gpurun_out
(timeout 900 python bench.py --steps 10 --warmup 3) > gpurun_out/r2_bench_ours_i.json 2> gpurun_out/r2_bench_ours_i.err; echo "bench ours rc=$?"; tail -3 gpurun_out/r2_bench_ours_i.err
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29642 bench.py --gpus 2 --steps 10 --warmup 3) > gpurun_out/r2_bench_2gpu_i.json 2> gpurun_out/r2_bench_2gpu_i.err; echo "bench N=2 rc=$?"; tail -2 gpurun_out/r2_bench_2gpu_i.err
python -c "
import json
for f in ('ours_i','2gpu_i'):
    d=json.loads(open('gpurun_out/r2_bench_%s.json'%f).read().strip().splitlines()[-1])
    r=d['roofline']
    print(f, d['n_gpus'], round(d['value'],2), 'TFLOP/s', round(d['ms_per_step'],2),'ms; e2e', round(d['e2e']['value'],2), round(d['e2e'].get('ms_per_step',0),2), d['e2e'].get('bit_identical_to_device_path'), 'roof', round(r['frac'],3), r.get('frac_of_sustained_peak'), r['kernel_ms'], r.get('frac_of_cublas_int8'), 'parity', (d.get('parity') or {}).get('max_ulp'), 'config4', round(d['config4']['value'],1))"
