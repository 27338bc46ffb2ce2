# round 2, call 50 (2 GPUs): config-4 shape (16384 columns, 4096 rows per rank here), broadcast panels: a product launch
# per panel against split-only panels with one product launch
mkdir -p gpurun_out
(timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29650 tools/sharded_probe.py 16384 4096 panels=1,2,4 nohost) 2>&1 | grep -E "max|product" | tee gpurun_out/r2_sharded_probe_2gpu_config4.txt
