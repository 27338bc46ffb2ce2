# round 2, call 17 (1 GPU): why are 2048^3 / 4096^3 slow under the new default choice (128-wide x 2 rounds, 240-wide x 4)?
mkdir -p gpurun_out
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv
for n in 2048 4096; do
  OZIMMU_B200_DEBUG=1 timeout 200 python tools/perf_probe.py $n 9 --iters 10 --shapes 00,p128,p240,p256,00,p192 --no-extras 2>&1 | tee -a gpurun_out/r2_perf_debug.txt
  OZIMMU_B200_LOCKSTEP=0 timeout 200 python tools/perf_probe.py $n 9 --iters 10 --shapes 00,p128,p240 --no-extras 2>&1 | sed "s/^/LOCKSTEP=0 /" | tee -a gpurun_out/r2_perf_debug.txt
done
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv
