# round 2, call 12 (1 GPU): short-product epilogue (pipelined TMEM loads, no release fence on the buffer hand-off) and
# the 128 x 128 tile (64 rows per CTA): parity tests of the touched paths, then timings at 512..8192
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_gemm.py tests/test_gpu_complex.py tests/test_gpu_batched.py -m gpu -q --maxfail=10) > gpurun_out/r2_t12.log 2>&1; echo "pytest gpu rc=$?"; tail -15 gpurun_out/r2_t12.log
for n in 512 1024 1536 2048; do
  timeout 200 python tools/perf_probe.py $n 9 --iters 20 --shapes 00,h128,p128,p192,p256 --no-extras --graph 2>&1 | tee -a gpurun_out/r2_perf_small_tiles.txt
done
timeout 200 python tools/perf_probe.py 4096 9 --iters 8 --shapes 00,p192,p256,00 --no-extras 2>&1 | tee -a gpurun_out/r2_perf_small_tiles.txt
timeout 200 python tools/perf_probe.py 8192 9 --iters 8 --shapes 00 --no-extras 2>&1 | tee -a gpurun_out/r2_perf_small_tiles.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_1024_b.csv python tools/perf_probe.py 1024 9 --iters 2 --shapes h128,p128 --no-extras > /dev/null 2>&1
python - <<PY
import csv
rows = [r for r in csv.reader(open('gpurun_out/r2_launches_1024_b.csv')) if len(r) > 5 and r[0].isdigit()]
for r in rows[-24:]:
    print(r[4][:90], r[-1], r[-2])
PY
