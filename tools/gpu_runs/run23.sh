mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_host_blocks.py "tests/test_gpu_gemm.py::test_gemm_host_equals_device_path" -x -q) > gpurun_out/t_blocks.log 2>&1; echo "pytest blocks rc=$?"; tail -15 gpurun_out/t_blocks.log
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()") 2>&1 | tail -2
(timeout 600 python tools/e2e_probe.py 8192) > gpurun_out/e2e_probe.log 2>&1; echo "probe rc=$?"; cat gpurun_out/e2e_probe.log
(timeout 600 python bench.py --steps 10 --warmup 3) > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "bench ours rc=$?"; python -c "
import json
d=json.loads(open('gpurun_out/bench_ours.json').read().strip().splitlines()[-1])
print('OURS', round(d['value'],2), 'TFLOP/s', round(d['ms_per_step'],2),'ms; e2e', round(d['e2e']['value'],2), d['e2e']['ms_per_step'], 'roof', d['roofline']['frac'], 'clocks', d['clocks'])"
