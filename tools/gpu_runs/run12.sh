mkdir -p gpurun_out
(timeout 300 python tests/golden/make_golden.py gpurun_out/golden) > gpurun_out/golden.log 2>&1; echo "golden rc=$?"; tail -3 gpurun_out/golden.log
(timeout 600 python -m pytest tests/test_gpu_complex.py -q --maxfail=5) > gpurun_out/t_cplx.log 2>&1; echo "complex rc=$?"; tail -12 gpurun_out/t_cplx.log
(timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_complex.py) > gpurun_out/t_all_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -4 gpurun_out/t_all_gpu.log
