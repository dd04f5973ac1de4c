mkdir -p gpurun_out
python -m pytest tests/ -q -m gpu -s > gpurun_out/gpu_tests.log 2>&1; tail -2 gpurun_out/gpu_tests.log; grep -E "teacher-forced|worst|golden rollout|resting|cooperative vs|Error|assert " gpurun_out/gpu_tests.log | head -30
python tools/quick_time.py 2>&1 | grep -v "Task id"
bash tools/gpu_timing.sh
