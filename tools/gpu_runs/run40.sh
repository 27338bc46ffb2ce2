(OZIMMU_B200_TEST_QUEUE=1 timeout 25 python -m pytest tests/test_gpu_queue.py -q -k "streamed or join") 2>&1 | tail -6
