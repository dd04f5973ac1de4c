python -m pytest tests/test_gpu_parity.py -q -m gpu -s 2>&1 | grep -E "worst|passed|failed|teacher"
python tools/step_timeline.py reach 8192 2>&1 | grep -v "Task id"
python tools/quick_time.py 2>&1 | grep -v "Task id"
bash tools/gpu_timing.sh
