mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "gripper or teacher" 2>&1 | tail -5
timeout 600 python tools/steady_time.py block_stack:2048 block_stack:256 block_rearrange:2048 2>&1 | grep -v "Task id" | tee gpurun_out/r2_11_timing.txt
