# round 2, call 6 (1 GPU): GPU suite after the pipelined one-pass strided split and the planes-first complex path; split ncu; ZGEMM timing
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q --maxfail=10) > gpurun_out/r2_t6.log 2>&1; echo "pytest gpu rc=$?"; tail -15 gpurun_out/r2_t6.log
(timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:split -c 6 --csv --log-file gpurun_out/r2_split_kernels_ncu.csv python tools/perf_probe.py 8192 9 --iters 1 --no-extras) > gpurun_out/r2_ncu_split.log 2>&1; echo "ncu rc=$?"; python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/r2_split_kernels_ncu.csv")) if len(r) > 10 and r[0].isdigit()]
for r in rows:
    print(r[0], r[4].split("(")[0][-40:], r[-3], r[-1])
PY
for z in 0 1; do OZIMMU_B200_ZGEMM_PLANES_FIRST=$z timeout 200 python tools/perf_probe.py 4096 9 --iters 8 --shapes 00 --zgemm --no-extras 2>&1 | grep ZGEMM; done | tee gpurun_out/r2_zgemm_4096.txt
for z in 0 1; do OZIMMU_B200_ZGEMM_PLANES_FIRST=$z timeout 200 python tools/perf_probe.py 2048 9 --iters 8 --shapes 00 --zgemm --no-extras 2>&1 | grep ZGEMM; done | tee -a gpurun_out/r2_zgemm_4096.txt
timeout 200 python tools/perf_probe.py 8192 9 --iters 6 2>&1 | grep -v "^$" | tee gpurun_out/r2_perf_8192.txt
