# round 2, call 28 (1 GPU): MMA-rate probe with RANDOM operand bytes (realistic toggle rate) and coalesced A reads,
# short (burst) and ~1.5 s per mode (sustained under the power cap)
mkdir -p gpurun_out
for n in 256 224 192; do
  timeout 60 tools/ubench/umma_rate_$n 20000 2>&1 | sed "s/^/burst /" | tee -a gpurun_out/r2_ubench_umma_rate_random.txt
  (nvidia-smi --query-gpu=clocks.sm,power.draw --format=csv,noheader -lms 250 > gpurun_out/r2_umma_clocks_$n.txt &) ; 
  timeout 200 tools/ubench/umma_rate_$n 4000000 2>&1 | sed "s/^/sustained /" | tee -a gpurun_out/r2_ubench_umma_rate_random.txt
  pkill -x nvidia-smi
  sort gpurun_out/r2_umma_clocks_$n.txt | uniq -c | sort -k1 -n -r | head -8
done
