# A/B: software-pipelined contact sweeps (Reach + one-block kernels) vs the plain loops (libpmg_x.so = -DPMG_SWEEP_PIPE=0)
mkdir -p gpurun_out
CASES="reach:8192 reach:1024 push:4096 pick_and_place:4096 slide:4096 push:512"
for i in 1 2; do
python tools/steady_time.py $CASES 2>&1 | grep "ms/step" | sed 's/^/pipe   /'
PMG_LIBRARY=$PWD/pybullet_multigoal_gym_b200/libpmg_x.so python tools/steady_time.py $CASES 2>&1 | grep "ms/step" | sed 's/^/plain  /'
done
timeout 600 python -m pytest tests/ -q -m gpu -x 2>&1 | tail -2
