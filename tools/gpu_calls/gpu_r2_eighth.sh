# Round 2, eighth GPU call (1 GPU): block_stack experiments -- sweep without the friction prefetch (x1), 2 envs per warp.
mkdir -p gpurun_out
{
echo "== base"; timeout 300 python tools/steady_time.py block_stack:2048 block_stack:1024 block_stack:256 2>&1 | grep -v "Task id"
echo "== base, 2 envs per warp"; PMG_COOP_EPB=2 timeout 300 python tools/steady_time.py block_stack:2048 block_stack:1024 2>&1 | grep -v "Task id"
export PMG_LIBRARY=$PWD/pybullet_multigoal_gym_b200/libpmg_x1.so
echo "== x1 (no friction prefetch)"; timeout 300 python tools/steady_time.py block_stack:2048 block_stack:1024 block_stack:256 block_rearrange:2048 2>&1 | grep -v "Task id"
echo "== x1, 2 envs per warp"; PMG_COOP_EPB=2 timeout 300 python tools/steady_time.py block_stack:2048 block_stack:1024 2>&1 | grep -v "Task id"
echo "== x1, 2 envs per warp, one-warp blocks"; PMG_COOP_WPB=1 PMG_COOP_EPB=2 timeout 300 python tools/steady_time.py block_stack:2048 2>&1 | grep -v "Task id"
echo "== x1, 1 env per warp"; PMG_COOP_EPB=1 timeout 300 python tools/steady_time.py block_stack:2048 block_stack:1024 2>&1 | grep -v "Task id"
} | tee gpurun_out/r2_eighth_timing.txt
