mkdir -p gpurun_out
python -m pytest tests/ -q -m gpu -s > gpurun_out/gpu_tests.log 2>&1; tail -3 gpurun_out/gpu_tests.log; grep -E "teacher-forced|worst|resting|Error|assert " gpurun_out/gpu_tests.log | head -20
echo "== coop block"; python tools/quick_time.py push:4096 pick_and_place:4096 push:512 pick_and_place:512 2>&1 | grep -v "Task id"
echo "== thread-per-env"; PMG_COOP_BLOCK=0 python tools/quick_time.py push:4096 pick_and_place:4096 push:512 pick_and_place:512 2>&1 | grep -v "Task id"
