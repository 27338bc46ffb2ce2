# round 2, call 26 (1 GPU): rate of the int8 CTA-pair MMA under SMEM traffic: SS vs TS mode, ring writes, A streamed
# through registers into TMEM (the shared-memory byte model of DESIGN 3.2 and whether TS mode would lift it)
mkdir -p gpurun_out
for n in 256 224 192 128; do
  timeout 60 tools/ubench/umma_rate_$n 20000 2>&1 | tee -a gpurun_out/r2_ubench_umma_rate.txt
done
