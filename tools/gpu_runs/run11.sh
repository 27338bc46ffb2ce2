mkdir -p gpurun_out
nvidia-smi -L
(timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/t_all_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -4 gpurun_out/t_all_gpu.log
for n in 1024 2048 4096 8192; do timeout 200 python tools/perf_probe.py $n 9 --shapes 00 --iters 10 --ref 2>&1 | grep -E "ozimmu_b200 n=|reference ozIMMU"; done 2>&1 | tee gpurun_out/sizes.log
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3) > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench2 rc=$?"; cat gpurun_out/bench_2gpu.json; tail -3 gpurun_out/bench_2gpu.err
(timeout 600 python bench.py --steps 10 --warmup 3) > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "bench ours rc=$?"; cat gpurun_out/bench_ours.json; tail -3 gpurun_out/bench_ours.err
