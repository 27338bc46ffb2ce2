# round 2, call 3 (2 GPUs): the C++ sharded path -- parity test, panel sweep (device + host operands), bench --gpus 2
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_dropin_cpp.py -m gpu -q -x) > gpurun_out/r2_t3.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/r2_t3.log
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/sharded_probe.py 8192) 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tee gpurun_out/r2_sharded_probe_2gpu.txt
(timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 5 --warmup 3) > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/r2_bench_2gpu.json; tail -5 gpurun_out/r2_bench_2gpu.err
