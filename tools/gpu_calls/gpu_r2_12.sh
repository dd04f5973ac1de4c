mkdir -p gpurun_out
{
echo "== stride 24 banks"; timeout 600 python tools/steady_time.py reach:8192 push:4096 pick_and_place:4096 slide:4096 block_stack:2048 reach:1024 2>&1 | grep -v "Task id"
export PMG_LIBRARY=$PWD/pybullet_multigoal_gym_b200/libpmg_plain.so
echo "== plain stride (sizeof)"; timeout 600 python tools/steady_time.py reach:8192 push:4096 pick_and_place:4096 slide:4096 block_stack:2048 2>&1 | grep -v "Task id"
} | tee gpurun_out/r2_12_timing.txt
