# round 2, call 52 (1 GPU): host-block / streamed-B tests after the removal, smoke
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_host_blocks.py tests/test_gpu_auto_and_dropin.py -m gpu -q --maxfail=5) > gpurun_out/r2_t52.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed" gpurun_out/r2_t52.log
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1
