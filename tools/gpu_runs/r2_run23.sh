# round 2, call 23 (1 GPU): L2 eviction-priority hints on the operand loads (low slices evict_last, high slices
# evict_first): 8192^3 time and DRAM bytes per setting
mkdir -p gpurun_out
for rep in 1 2; do
for hot in "0,0" "3,2" "2,2" "4,3" "1,1" "5,4" "3,3"; do
  OZIMMU_B200_L2_HOT=$hot timeout 200 python tools/perf_probe.py 8192 9 --iters 8 --shapes 00 --no-extras 2>&1 | sed "s/^/L2_HOT=$hot /" | tee -a gpurun_out/r2_l2_hints.txt
done
done
for hot in "0,0" "3,2" "4,3" "2,2"; do
  OZIMMU_B200_L2_HOT=$hot timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:oz_gemm_pair -c 1 --csv --log-file gpurun_out/r2_l2_hints_ncu_$hot.csv python tools/perf_probe.py 8192 9 --iters 1 --shapes 00 --no-extras > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2_l2_hints_ncu_$hot.csv')) if len(r)>5 and r[0].isdigit()]
print('L2_HOT=$hot', [(r[-3], r[-1], r[-2]) for r in rows])
PY
done 2>&1 | tee -a gpurun_out/r2_l2_hints.txt
