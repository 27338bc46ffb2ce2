mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()"
for rep in 1 2; do
for ns in 0 32 100 500; do
  echo -n "rep$rep IDLE_SLEEP=$ns: "
  OZIMMU_B200_IDLE_SLEEP=$ns timeout 200 python tools/perf_probe.py 8192 9 --iters 20 2>&1 | head -1
done
done 2>&1 | tee gpurun_out/sweep6.log
