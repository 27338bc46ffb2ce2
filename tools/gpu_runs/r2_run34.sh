# round 2, call 34 (1 GPU): after allocating the pacing counters in one piece: GPU tests of the product path, then ours
# vs the reference vs cuBLAS DGEMM at 1024 .. 8192
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_kernels.py tests/test_gpu_host_blocks.py -m gpu -q --maxfail=10) > gpurun_out/r2_t34.log 2>&1; echo "pytest gpu rc=$?"; grep -E "passed|failed" gpurun_out/r2_t34.log
rm -f gpurun_out/r2_sizes_vs_reference.txt
for n in 1024 1536 2048 3072 4096 6144 8192; do
  timeout 300 python tools/perf_probe.py $n 9 --ref --iters 10 2>&1 | grep -E "^ozimmu_b200 n=|cuBLAS DGEMM|reference ozIMMU" | tee -a gpurun_out/r2_sizes_vs_reference.txt
done
