mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" 
for rep in 1 2; do
for cfg in "0 2" "0 0" "16 2" "0 1" "0 4" "4 2"; do
  set -- $cfg
  echo -n "rep$rep PREFETCH=$1 LOCKSTEP=$2: "
  OZIMMU_B200_PREFETCH=$1 OZIMMU_B200_LOCKSTEP=$2 timeout 200 python tools/perf_probe.py 8192 9 --iters 20 2>&1 | head -1
done
done 2>&1 | tee gpurun_out/sweep5.log
for cfg in "0 2" "0 0"; do
  set -- $cfg
  OZIMMU_B200_PREFETCH=$1 OZIMMU_B200_LOCKSTEP=$2 timeout 300 ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,sm__cycles_elapsed.avg.per_second --clock-control none -k regex:oz_gemm_pair -c 1 python tools/perf_probe.py 8192 9 --iters 1 2>&1 | grep -E "dram__|gpu__time|hit_rate|per_second" | sed "s/^/[$1 $2] /"
done 2>&1 | tee gpurun_out/sweep5_ncu.log
