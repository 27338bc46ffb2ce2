# round 2, call 14 (1 GPU): full GPU suite on the current tree; tail-split probe at 4096 / 6144 / 2048; small sizes per
# tile shape; ncu of the 128 x 128 tile at 1024^3; compute-sanitizer memcheck of the new tile on a small problem
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q --maxfail=10) > gpurun_out/r2_t14.log 2>&1; echo "pytest gpu rc=$?"; tail -6 gpurun_out/r2_t14.log
timeout 300 python tools/tail_split_probe.py 4096 2>&1 | tee gpurun_out/r2_tail_split_probe.txt
timeout 300 python tools/tail_split_probe.py 6144 2>&1 | tee -a gpurun_out/r2_tail_split_probe.txt
timeout 300 python tools/tail_split_probe.py 2048 2>&1 | tee -a gpurun_out/r2_tail_split_probe.txt
for n in 1536 2048 3072; do
  timeout 200 python tools/perf_probe.py $n 9 --iters 20 --shapes 00,h128,p128,p192,p240,p256 --no-extras 2>&1 | tee -a gpurun_out/r2_perf_small_tiles_b.txt
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:oz_gemm_pair_kernel -s 2 -c 1 -f -o gpurun_out/r2_prof_h128_1024 python tools/perf_probe.py 1024 9 --iters 2 --shapes h128 --no-extras > gpurun_out/r2_ncu_h128.log 2>&1; echo "ncu rc=$?"
OZ_SAN=1 timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_gemm.py -m gpu -q -k "cluster_shapes_same_bits and 64" > gpurun_out/r2_sanitizer_h128.txt 2>&1; echo "sanitizer rc=$?"; tail -5 gpurun_out/r2_sanitizer_h128.txt
