# round 2, call 27 (1 GPU): the same MMA-rate probe over ~1 s per mode: what remains of the TS-mode advantage under the
# power cap (sustained TOP/s from event times)
mkdir -p gpurun_out
for n in 256 192; do
  timeout 120 tools/ubench/umma_rate_$n 3000000 2>&1 | tee -a gpurun_out/r2_ubench_umma_rate_sustained.txt
done
nvidia-smi --query-gpu=clocks.sm,power.draw,power.limit,clocks_event_reasons.active --format=csv
