export PMG_LIBRARY=pybullet_multigoal_gym_b200/libpmg_timing.so
for t in push:512 pick_and_place:512 push:4096; do echo "== $t"; python tools/coop_timing.py $t 2>&1 | grep -v "Task id"; done
