python -m pytest tests/test_gpu_parity.py -q -m gpu -k "reach" -s 2>&1 | grep -E "worst|passed|failed"
python tools/step_timeline.py reach 8192 2>&1 | grep -v "Task id"
python tools/quick_time.py reach:8192 reach:65536 2>&1 | grep -v "Task id"
