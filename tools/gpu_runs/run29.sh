python -c "import torch; torch.zeros(1).cuda()"
for rep in 1 2; do
for cfg in "0 2" "8 2" "16 2" "32 2" "0 1" "0 4" "0 0" "16 4"; do
  set -- $cfg
  echo -n "rep$rep PREFETCH=$1 LOCKSTEP=$2: "
  OZIMMU_B200_PREFETCH=$1 OZIMMU_B200_LOCKSTEP=$2 timeout 200 python tools/perf_probe.py 8192 9 --iters 20 2>&1 | head -1
done
done 2>&1 | tee gpurun_out/sweep7.log
