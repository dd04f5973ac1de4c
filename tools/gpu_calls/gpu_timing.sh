PMG_LIBRARY=pybullet_multigoal_gym_b200/libpmg_timing.so python tools/coop_timing.py 2>&1 | grep -v "Task id"
