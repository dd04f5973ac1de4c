{
echo "== SPTS 12 (default)"; timeout 600 python tools/steady_time.py push:4096 pick_and_place:4096 slide:4096 2>&1 | grep -v "Task id"
export PMG_LIBRARY=$PWD/pybullet_multigoal_gym_b200/libpmg_spts11.so
echo "== SPTS 11 (two-warp lockstep blocks fit)"; timeout 600 python tools/steady_time.py push:4096 pick_and_place:4096 slide:4096 2>&1 | grep -v "Task id"
echo "== SPTS 11, one-warp blocks"; PMG_COOP_WPB=1 timeout 600 python tools/steady_time.py push:4096 pick_and_place:4096 2>&1 | grep -v "Task id"
} | tee gpurun_out/r2_17_timing.txt
