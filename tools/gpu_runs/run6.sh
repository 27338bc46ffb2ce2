mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_gemm.py -q --maxfail=6) > gpurun_out/t_gemm.log 2>&1; echo "gemm rc=$?"; tail -3 gpurun_out/t_gemm.log
for cfg in "0 0" "16 0" "0 2" "16 2" "16 1" "16 4" "32 2" "8 2"; do
  set -- $cfg
  echo "== PREFETCH=$1 LOCKSTEP=$2"
  OZIMMU_B200_PREFETCH=$1 OZIMMU_B200_LOCKSTEP=$2 timeout 200 python tools/perf_probe.py 8192 9 --shapes p192 --iters 5 2>&1 | head -1
  OZIMMU_B200_PREFETCH=$1 OZIMMU_B200_LOCKSTEP=$2 timeout 200 python tools/perf_probe.py 4096 9 --shapes p192 --iters 10 2>&1 | head -1
done 2>&1 | tee gpurun_out/sweep1.log
for cfg in "0 0" "16 2"; do
  set -- $cfg
  OZIMMU_B200_PREFETCH=$1 OZIMMU_B200_LOCKSTEP=$2 timeout 300 ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,sm__cycles_elapsed.avg.per_second,sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:oz_gemm_pair -c 1 python tools/perf_probe.py 8192 9 --shapes p192 --iters 1 2>&1 | grep -E "dram__|gpu__time|hit_rate|per_second|imma" | sed "s/^/[$1 $2] /"
done 2>&1 | tee gpurun_out/sweep1_ncu.log
