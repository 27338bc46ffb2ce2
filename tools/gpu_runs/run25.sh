mkdir -p gpurun_out
: > gpurun_out/prefetch_coop.log
for rep in 1 2; do
for cfg in "0 1" "8 1" "16 1" "32 1" "48 1" "16 0"; do
  set -- $cfg
  echo -n "rep$rep PREFETCH=$1 COOP=$2: " >> gpurun_out/prefetch_coop.log
  OZIMMU_B200_PREFETCH=$1 OZIMMU_B200_PREFETCH_COOP=$2 timeout 300 python tools/perf_probe.py 8192 9 --iters 10 2>&1 | grep "^ozimmu_b200" >> gpurun_out/prefetch_coop.log
done
done
cat gpurun_out/prefetch_coop.log
