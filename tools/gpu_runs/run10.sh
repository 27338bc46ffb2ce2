mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_kernels.py -q --maxfail=4 -k "pair_product") > gpurun_out/t_kernels_pair.log 2>&1; echo "pair rc=$?"; tail -3 gpurun_out/t_kernels_pair.log
(timeout 300 python -m pytest tests/test_gpu_gemm.py -q --maxfail=6 -k "cluster_shapes or oracle") > gpurun_out/t_gemm.log 2>&1; echo "gemm rc=$?"; tail -3 gpurun_out/t_gemm.log
for sh in p192 p256; do
  timeout 200 python tools/perf_probe.py 8192 9 --shapes $sh --iters 10 2>&1 | grep -E "ozimmu_b200 n="
  timeout 200 python tools/perf_probe.py 4096 9 --shapes $sh --iters 10 2>&1 | grep -E "ozimmu_b200 n="
  timeout 200 python tools/perf_probe.py 16384 9 --shapes $sh --iters 3 2>&1 | grep -E "ozimmu_b200 n="
done 2>&1 | tee gpurun_out/sweep3.log
for sh in p192 p256; do
timeout 300 ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,sm__cycles_elapsed.avg.per_second,sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active,l1tex__m_xbar2l1tex_read_bytes.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:oz_gemm_pair -c 1 python tools/perf_probe.py 8192 9 --shapes $sh --iters 1 2>&1 | grep -E "dram__|gpu__time|hit_rate|per_second|imma|xbar|lts__thr" | sed "s/^/[$sh] /"
done 2>&1 | tee gpurun_out/sweep3_ncu.log
