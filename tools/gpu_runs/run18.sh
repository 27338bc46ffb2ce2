mkdir -p gpurun_out
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:oz_gemm_pair -c 1 -o gpurun_out/prof_pair256_r1e python tools/perf_probe.py 8192 9 --iters 1) > gpurun_out/ncu5.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu5.log
(timeout 600 python bench.py --impl reference --steps 10 --warmup 3) > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"; python -c "
import json
d=json.loads(open('gpurun_out/bench_ref.json').read().strip().splitlines()[-1])
print('REF', round(d['value'],2), 'TFLOP/s', round(d['ms_per_step'],2),'ms; e2e', round(d['e2e']['value'],2), 'clocks', d['clocks'])"
(timeout 600 python bench.py --steps 10 --warmup 3) > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "bench ours rc=$?"; python -c "
import json
d=json.loads(open('gpurun_out/bench_ours.json').read().strip().splitlines()[-1])
print('OURS', round(d['value'],2), 'TFLOP/s', round(d['ms_per_step'],2),'ms; e2e', round(d['e2e']['value'],2), 'roof', d['roofline'], 'clocks', d['clocks'], 'launches', d['gpu_launches'])"; tail -2 gpurun_out/bench_ours.err
(timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3) > gpurun_out/bench_ncu.log 2>&1; echo "ncu launches rc=$?"
