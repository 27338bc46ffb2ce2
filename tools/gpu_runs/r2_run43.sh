# round 2, call 43 (2 GPUs): why is the receiving rank of the weak-scaling step slower than the root by more than the
# broadcast?  sharded probe (per-rank times), default vs no lockstep
mkdir -p gpurun_out
for ls in 2 0; do
  (OZIMMU_B200_LOCKSTEP=$ls timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2965$ls tools/sharded_probe.py 8192 8192 panels=1 nohost) 2>&1 | grep -E "max|product" | sed "s/^/LOCKSTEP=$ls /" | tee -a gpurun_out/r2_sharded_probe_2gpu_b.txt
done
nvidia-smi --query-gpu=index,clocks.sm,power.draw,temperature.gpu --format=csv
