# round 2, call 16 (1 GPU): k-aware tile cost model -- host-operand entry (its block launches now choose by SM time per
# area: 256-wide at k = 8192 instead of 128-wide) and the default choice at 1024..8192
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_host_blocks.py tests/test_gpu_batched.py -m gpu -q --maxfail=10) > gpurun_out/r2_t16.log 2>&1; echo "pytest gpu rc=$?"; tail -4 gpurun_out/r2_t16.log
for w in 0 128 0 256 192; do
  OZIMMU_B200_TILE_N=$w timeout 200 python tools/e2e_probe.py 8192 768:768 1024:1024 2>&1 | grep gemm_host | sed "s/^/TILE_N=$w /" | tee -a gpurun_out/r2_e2e_tile_width.txt
done
for n in 1024 1536 2048 3072 4096 6144 8192; do
  timeout 200 python tools/perf_probe.py $n 9 --iters 10 --shapes 00 --no-extras 2>&1 | tee -a gpurun_out/r2_perf_default_dispatch.txt
done
