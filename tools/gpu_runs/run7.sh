mkdir -p gpurun_out
export OZIMMU_B200_DEBUG=1
(timeout 300 python -m pytest tests/test_gpu_kernels.py -q --maxfail=4 -k "pair_product") > gpurun_out/t_kernels_pair.log 2>&1; echo "pair rc=$?"; tail -4 gpurun_out/t_kernels_pair.log
(timeout 300 python -m pytest tests/test_gpu_gemm.py -q --maxfail=6 -k "cluster_shapes or oracle") > gpurun_out/t_gemm.log 2>&1; echo "gemm rc=$?"; tail -4 gpurun_out/t_gemm.log
for sh in p192 c21 c12 c22; do
  timeout 200 python tools/perf_probe.py 8192 9 --shapes $sh --iters 5 2>&1 | grep -E "resident|ozimmu_b200 n="
  timeout 200 python tools/perf_probe.py 4096 9 --shapes $sh --iters 10 2>&1 | grep -E "ozimmu_b200 n="
done 2>&1 | tee gpurun_out/sweep2.log
