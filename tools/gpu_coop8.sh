mkdir -p gpurun_out
python -m pytest tests/ -q -m gpu -x -s > gpurun_out/gpu_tests.log 2>&1; tail -2 gpurun_out/gpu_tests.log; grep -E "teacher-forced|worst|golden rollout|resting" gpurun_out/gpu_tests.log
python tools/step_timeline.py reach 8192 2>&1 | grep -v "Task id"
python tools/quick_time.py 2>&1 | grep -v "Task id"
bash tools/gpu_timing.sh
