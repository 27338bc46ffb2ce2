# round 2, call 46 (1 GPU): compute-sanitizer memcheck over the paths touched this round (final code)
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_complex.py tests/test_gpu_host_blocks.py tests/test_gpu_auto_and_dropin.py -m gpu -q -x -k "cluster_shapes or forced_tiles or fused_blocks_equal_whole or block_schedules or graph_capture or odd or ragged" > gpurun_out/r2_sanitizer_final.txt 2>&1; echo "sanitizer rc=$?"; tail -6 gpurun_out/r2_sanitizer_final.txt
