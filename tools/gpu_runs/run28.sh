mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29557 tools/ubench/sharded_breakdown.py 2>&1 | grep "^world" | tee gpurun_out/sharded_breakdown.log
(timeout 600 python -m pytest tests/test_gpu_auto_and_dropin.py tests/test_gpu_complex.py -x -q -k "sgemm or dropin or passthrough or complex") 2>&1 | tail -5
