mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 || { echo "SMOKE FAILED/HUNG"; exit 1; }
timeout 1200 python -m pytest tests/ -q -m gpu -x 2>&1 | tail -6
timeout 600 python tools/steady_time.py reach:8192 push:4096 block_stack:2048 2>&1 | grep -v "Task id"
