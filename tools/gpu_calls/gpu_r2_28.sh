# per-phase cycles with the sweep-internal counters (timing build)
export PMG_LIBRARY=$PWD/pybullet_multigoal_gym_b200/libpmg_timing.so
for a in "reach:1024" "down reach:1024" "reach:8192" "down reach:8192" "push:512" "push:4096"; do echo "== $a"; python tools/coop_timing.py $a 2>&1 | grep -v "^Task id"; done
