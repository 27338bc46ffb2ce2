mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_host_blocks.py -x -q) 2>&1 | tail -3
(timeout 300 python tools/e2e_probe.py 8192 768:768:0 768:768:1 1024:1024:0 1024:1024:1 768:768:0 768:768:1 1024:1024:1 1536:1536:1) 2>&1 | tee gpurun_out/e2e_probe_taper.log | head -9
