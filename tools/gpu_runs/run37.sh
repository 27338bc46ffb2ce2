mkdir -p gpurun_out
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()") 2>&1 | tail -1
(timeout 400 python bench.py --steps 5 --warmup 3) > gpurun_out/bench_check.json 2> gpurun_out/bench_check.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/bench_check.json; tail -3 gpurun_out/bench_check.err
