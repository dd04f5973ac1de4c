# A/B: pipelined Reach contact sweep (cross terms of consecutive rows) vs the plain sweep (libpmg_x.so = -DPMG_SWEEP_PIPE=0)
mkdir -p gpurun_out
for i in 1 2; do
python tools/steady_time.py reach:8192 reach:1024 2>&1 | grep "ms/step" | sed 's/^/pipe   /'
PMG_LIBRARY=$PWD/pybullet_multigoal_gym_b200/libpmg_x.so python tools/steady_time.py reach:8192 reach:1024 2>&1 | grep "ms/step" | sed 's/^/plain  /'
done
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "reach" 2>&1 | tail -2
