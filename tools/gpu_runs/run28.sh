mkdir -p gpurun_out
(timeout 1800 python -m pytest tests -m gpu -x -q) > gpurun_out/t_all_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -4 gpurun_out/t_all_gpu.log
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()") 2>&1 | tail -2
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:oz_gemm_pair -c 1 -o gpurun_out/prof_pair256_r1f python tools/perf_probe.py 8192 9 --iters 1) > gpurun_out/ncu6.log 2>&1; echo "ncu rc=$?"
(timeout 600 python bench.py --steps 10 --warmup 3) > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "bench ours rc=$?"; python -c "
import json
d=json.loads(open('gpurun_out/bench_ours.json').read().strip().splitlines()[-1])
print('OURS', round(d['value'],2), 'TFLOP/s', round(d['ms_per_step'],2),'ms; e2e', round(d['e2e']['value'],2), 'roof', d['roofline']['frac'], 'clocks', d['clocks'])"; tail -2 gpurun_out/bench_ours.err
