(OZIMMU_B200_TEST_QUEUE=1 timeout 45 python -m pytest tests/test_gpu_queue.py -q) 2>&1 | tail -12
