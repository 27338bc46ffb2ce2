# round 2, call 38 (1 GPU): MMA-rate probe for UMMA M = 128 (64 rows per CTA), SS mode, with and without the operand ring
mkdir -p gpurun_out
for b in umma_rate_m128_128 umma_rate_m128_256 umma_rate_128; do
  timeout 30 tools/ubench/$b 20000 2>&1 | tee -a gpurun_out/r2_ubench_umma_rate_m128.txt
done
