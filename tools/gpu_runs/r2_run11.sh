# round 2, call 11 (1 GPU): ncu --set full of the product kernel at 1024^3 and 2048^3 (where does a short product's time go)
mkdir -p gpurun_out
for n in 1024 2048; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:oz_gemm_pair_kernel -s 2 -c 1 -f -o gpurun_out/r2_prof_small_$n python tools/perf_probe.py $n 9 --iters 2 --shapes 00 --no-extras > gpurun_out/r2_ncu_small_$n.log 2>&1; echo "ncu $n rc=$?"
done
ls -la gpurun_out/*.ncu-rep
