# round 2, call 33 (1 GPU): ours vs the reference vs cuBLAS DGEMM at 1024 .. 8192 (final code)
mkdir -p gpurun_out
python -c "import torch; a=torch.rand(4096,4096,device='cuda'); [a@a for _ in range(200)]; torch.cuda.synchronize()"   # clocks up
for n in 1024 1536 2048 3072 4096 6144 8192; do
  timeout 300 python tools/perf_probe.py $n 9 --ref --iters 10 2>&1 | grep -E "^ozimmu_b200 n=|cuBLAS DGEMM|cuBLAS int8|reference ozIMMU" | tee -a gpurun_out/r2_sizes_vs_reference.txt
done
