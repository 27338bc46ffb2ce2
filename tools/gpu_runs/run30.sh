mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29557 tools/ubench/sharded_breakdown.py 2>&1 | grep "^world\|Error\|error" | tee gpurun_out/sharded_breakdown4.log
