# round 2, call 13 (4 GPUs): host <-> device copy rates with 1 / 2 / 4 GPUs active at once (the floor under the sharded
# host-operand entry), and the sharded step at the config-4 shape (16384 columns, 4096 rows per rank): one panel against
# panels that trigger their split only (one product launch) or a product launch each
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo.txt 2>&1; head -20 gpurun_out/r2_topo.txt
(timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29611 tools/ubench/pcie_multi.py) 2>&1 | grep -v "^W\|^\*\*\*" | tee gpurun_out/r2_ubench_pcie_multi_4gpu.txt
(timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29612 tools/ubench/pcie_multi.py --no-bind) 2>&1 | grep -v "^W\|^\*\*\*" | tee gpurun_out/r2_ubench_pcie_multi_4gpu_nobind.txt
(timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29613 tools/sharded_probe.py 16384 4096 panels=1,2,4 nohost) 2>&1 | grep -v "^W\|^\*\*\*" | tee gpurun_out/r2_sharded_probe_4gpu_config4.txt
