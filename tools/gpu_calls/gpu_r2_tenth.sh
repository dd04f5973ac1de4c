# Round 2, tenth GPU call (1 GPU): gripper-base collisions: parity suite, timing of the multi-block tasks.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -q -m gpu -x 2>&1 | tail -25 | tee gpurun_out/r2_tenth_tests.txt
timeout 600 python tools/steady_time.py block_stack:2048 block_stack:256 block_rearrange:2048 reach:8192 2>&1 | grep -v "Task id" | tee gpurun_out/r2_tenth_timing.txt
