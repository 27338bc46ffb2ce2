mkdir -p gpurun_out
for n in 8 4 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n tools/config4.py 16384 3 2>/dev/null | grep config
done | tee gpurun_out/config4.jsonl
timeout 600 python tools/config4.py 16384 2 2>/dev/null | grep config | tee -a gpurun_out/config4.jsonl
