mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_sharded.py -x -q) 2>&1 | tail -12
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29557 tools/ubench/sharded_breakdown.py 2>&1 | grep "^world\|Error\|error" | tee gpurun_out/sharded_breakdown2.log
