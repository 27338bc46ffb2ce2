mkdir -p gpurun_out
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()") 2>&1 | tail -3
(timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/t_all_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -4 gpurun_out/t_all_gpu.log
(timeout 600 python bench.py --impl reference --steps 10 --warmup 3) > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"; cat gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err
(timeout 600 python bench.py --steps 10 --warmup 3) > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "bench ours rc=$?"; cat gpurun_out/bench_ours.json; tail -3 gpurun_out/bench_ours.err
(timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3) > gpurun_out/bench_ncu.log 2>&1; echo "ncu launches rc=$?"
