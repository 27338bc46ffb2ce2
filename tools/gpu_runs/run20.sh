mkdir -p gpurun_out
for cfg in "16 2" "0 2" "8 2" "32 2" "16 1" "16 3" "16 4" "0 0"; do
  set -- $cfg
  echo -n "PREFETCH=$1 LOCKSTEP=$2: "
  OZIMMU_B200_PREFETCH=$1 OZIMMU_B200_LOCKSTEP=$2 timeout 200 python tools/perf_probe.py 8192 9 --iters 10 2>&1 | head -1
done 2>&1 | tee gpurun_out/sweep4.log
