timeout 900 python -m pytest tests/test_gpu_api.py tests/test_gpu_sharded.py -q -m gpu -x 2>&1 | tail -5
