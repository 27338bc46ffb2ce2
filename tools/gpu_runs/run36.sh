(timeout 600 python -m pytest tests/test_gpu_batched.py tests/test_gpu_auto_and_dropin.py tests/test_gpu_dropin_cpp.py -x -q) 2>&1 | tail -8
