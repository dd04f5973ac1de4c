export PMG_LIBRARY=$PWD/pybullet_multigoal_gym_b200/libpmg_timing.so
for t in reach:4 reach:592 reach:8192; do echo "== $t"; timeout 300 python tools/coop_timing.py $t 2>&1 | grep -v "Task id"; echo "== $t down"; timeout 300 python tools/coop_timing.py down $t 2>&1 | grep -v "Task id"; done | tee gpurun_out/r2_18_lone_warp_cycles.txt
