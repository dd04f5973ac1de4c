timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -s -k "closed_loop" 2>&1 | tail -8
