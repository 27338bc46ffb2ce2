# round 2, call 4 (1 GPU): reference-driver modes, multi-launch cost in isolation, e2e trace, config-3 sweep with the shipped kernel
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_reference_driver.py -m gpu -q -x -k matfile -s) > gpurun_out/r2_t4.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2_t4.log
timeout 300 python tools/streamed_probe.py 8192 2>&1 | tee gpurun_out/r2_streamed_probe.txt
OZIMMU_B200_E2E_TRACE=1 timeout 300 python - <<'PY' 2>&1 | tail -80 | tee gpurun_out/r2_e2e_trace.txt
import torch, ozimmu_b200 as oz
n = 8192
a = torch.rand(n * n, dtype=torch.float64).pin_memory(); b = torch.rand(n * n, dtype=torch.float64).pin_memory()
c = torch.zeros(n * n, dtype=torch.float64).pin_memory()
h = oz.create()
import os
os.environ["OZIMMU_B200_E2E_TRACE"] = "0"
for _ in range(3):
    oz.gemm_host(h, 0, 0, n, n, n, 1.0, a, n, b, n, 0.0, c, n, oz.fp64_int8(9))
os.environ["OZIMMU_B200_E2E_TRACE"] = "1"
oz.gemm_host(h, 0, 0, n, n, n, 1.0, a, n, b, n, 0.0, c, n, oz.fp64_int8(9))
PY
timeout 900 python tools/sweeps.py split 8192 2>&1 | tee gpurun_out/r2_config3_split_sweep_8192.csv | tail -40
