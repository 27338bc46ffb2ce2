mkdir -p gpurun_out
(time timeout 1700 python -m pytest tests -m gpu -x -q) > gpurun_out/t_all_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -6 gpurun_out/t_all_gpu.log
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()") 2>&1 | tail -2
(timeout 600 python bench.py --impl reference --steps 10 --warmup 3) > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "bench ref rc=$?"
(timeout 600 python bench.py --steps 10 --warmup 3) > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "bench ours rc=$?"
python -c "
import json
for f in ('ours','reference'):
    d=json.loads(open('gpurun_out/bench_%s.json'%f).read().strip().splitlines()[-1])
    print(f, round(d['value'],2), 'TFLOP/s', round(d['ms_per_step'],2),'ms; e2e', round(d['e2e']['value'],2), d['e2e'].get('ms_per_step'), 'roof', (d.get('roofline') or {}).get('frac'), 'launches', d.get('gpu_launches'), 'clocks', d['clocks'])"
(timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1) > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu launches rc=$?"; wc -l gpurun_out/launches.csv
(timeout 900 ncu --set full --clock-control none --import-source on -k regex:oz_gemm_pair -c 1 -o gpurun_out/prof_pair256_r1g python tools/perf_probe.py 8192 9 --iters 1) > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"; ls -la gpurun_out/*.ncu-rep
