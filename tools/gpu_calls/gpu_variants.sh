mkdir -p gpurun_out
python -m pytest tests/ -q -m gpu -x -s > gpurun_out/gpu_tests.log 2>&1; tail -3 gpurun_out/gpu_tests.log; grep -E "teacher-forced|worst|golden rollout" gpurun_out/gpu_tests.log
python tools/quick_time.py reach:8192 push:4096 block_stack:2048 2>&1 | grep -v "Task id"
timeout 300 compute-sanitizer --tool racecheck --racecheck-report analysis python tools/prof_one.py reach 16 11 down > gpurun_out/racecheck_coop.log 2>&1; tail -4 gpurun_out/racecheck_coop.log
timeout 200 compute-sanitizer --tool memcheck python tools/prof_one.py reach 16 11 down > gpurun_out/memcheck_coop.log 2>&1; tail -2 gpurun_out/memcheck_coop.log
