# First GPU run of the opt-in cooperative multi-block kernel (PMG_COOP_STACK=1): the block_stack / block_rearrange
# parity tests through it, then timing against the thread-per-env kernel.
mkdir -p gpurun_out
PMG_COOP_STACK=1 python -m pytest tests/ -q -m gpu -s -k "block_stack or rearrange or variant or curriculum or decomposition" > gpurun_out/gpu_tests_coop_stack.log 2>&1; tail -3 gpurun_out/gpu_tests_coop_stack.log
echo "== cooperative"; PMG_COOP_STACK=1 python tools/quick_time.py block_stack:2048 block_stack:256 2>&1 | grep -v "Task id"
echo "== thread-per-env"; python tools/quick_time.py block_stack:2048 block_stack:256 2>&1 | grep -v "Task id"
