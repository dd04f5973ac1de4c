mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -q -m gpu -x 2>&1 | tail -6
timeout 900 python tools/her_reset_timing.py 2>&1 | grep -v "Task id" | tee gpurun_out/r2_her_reset_timing.txt
timeout 300 python tools/steady_time.py reach:8192 push:4096 2>&1 | grep -v "Task id"
