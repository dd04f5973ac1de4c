# Round 2, seventh GPU call (1 GPU): static-box narrowphase fast path + lockstep multi-warp blocks: parity suite, then A/B timing.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -q -m gpu -x 2>&1 | tail -15 | tee gpurun_out/r2_seventh_tests.txt
{
echo "== default geometry"; timeout 600 python tools/steady_time.py reach:8192 push:4096 pick_and_place:4096 slide:4096 block_stack:2048 block_stack:1024 block_stack:256 reach:1024 2>&1 | grep -v "Task id"
echo "== one-warp blocks"; PMG_COOP_WPB=1 timeout 600 python tools/steady_time.py reach:8192 block_stack:2048 block_stack:1024 2>&1 | grep -v "Task id"
echo "== two-warp blocks"; PMG_COOP_WPB=2 timeout 600 python tools/steady_time.py block_stack:2048 block_stack:1024 push:2048 pick_and_place:2048 2>&1 | grep -v "Task id"
echo "== one-warp blocks"; PMG_COOP_WPB=1 timeout 600 python tools/steady_time.py push:2048 pick_and_place:2048 2>&1 | grep -v "Task id"
} | tee gpurun_out/r2_seventh_timing.txt
