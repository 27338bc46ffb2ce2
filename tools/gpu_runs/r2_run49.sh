# round 2, call 49 (8 GPUs): host <-> device copy rates with 1 / 2 / 4 / 8 GPUs active at once (the floor under the 8-GPU
# host-operand step)
mkdir -p gpurun_out
(timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29649 tools/ubench/pcie_multi.py) 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tee gpurun_out/r2_ubench_pcie_multi_8gpu.txt
