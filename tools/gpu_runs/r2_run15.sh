# round 2, call 15 (1 GPU): per-round tile cost by tile shape and k (for the dispatch cost model), whole-problem check at
# 4096^3 / 8192^3 incl. 128-wide tiles, compute-sanitizer memcheck of the 128 x 128 tile
mkdir -p gpurun_out
timeout 600 python tools/tile_cost_probe.py 2>&1 | tee gpurun_out/r2_tile_cost_probe.txt
timeout 200 python tools/perf_probe.py 4096 9 --iters 8 --shapes 00,p128,p192,p256,p128,p192 --no-extras 2>&1 | tee gpurun_out/r2_perf_widths_4096_8192.txt
timeout 300 python tools/perf_probe.py 8192 9 --iters 6 --shapes 00,p128,p256,p128 --no-extras 2>&1 | tee -a gpurun_out/r2_perf_widths_4096_8192.txt
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_complex.py -m gpu -q -k "cluster_shapes_same_bits or forced_tiles" > gpurun_out/r2_sanitizer_h128.txt 2>&1; echo "sanitizer rc=$?"; tail -5 gpurun_out/r2_sanitizer_h128.txt
