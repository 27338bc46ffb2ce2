# round 2, call 5 (1 GPU): whole GPU suite after the one-pass strided split, reference-driver modes, PCIe 2-D copy rates,
# ncu of the split kernels
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q --maxfail=10) > gpurun_out/r2_t5.log 2>&1; echo "pytest gpu rc=$?"; tail -15 gpurun_out/r2_t5.log
timeout 300 python tools/ubench/pcie_2d.py 2>&1 | tee gpurun_out/r2_ubench_pcie_2d.txt
(timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:split -c 12 --csv --log-file gpurun_out/r2_split_kernels_ncu.csv python tools/perf_probe.py 8192 9 --iters 1 --no-extras) > gpurun_out/r2_ncu_split.log 2>&1; echo "ncu rc=$?"; cat gpurun_out/r2_split_kernels_ncu.csv | tail -40
timeout 200 python tools/perf_probe.py 8192 9 --iters 8 2>&1 | tee gpurun_out/r2_perf_8192.txt
