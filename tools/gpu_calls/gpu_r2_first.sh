# Round 2, first GPU call: (1) is a real pybullet/gym or baseline/_ref on the box? (2) the cooperative multi-block
# kernel (PMG_COOP_STACK=1) through the block_stack / block_rearrange parity tests, racecheck, timing vs thread-per-env.
mkdir -p gpurun_out
{
  echo "== probe for the real reference backend"; 
  python -c "import pybullet, gym; print('pybullet', pybullet.__file__, 'gym', gym.__file__)" 2>&1 | tail -1
  ls -la baseline/_ref 2>&1 | head -5
  python -m pip download pybullet==3.0.6 --no-deps -d /tmp/pb 2>&1 | tail -2
  find / -iname '*pybullet*' -not -path '*/proc/*' -not -path "$GRAFT_REPO_ROOT/*" 2>/dev/null | head -5
  nproc; lscpu | grep 'Model name'
} > gpurun_out/r02_pybullet_probe.txt 2>&1
cat gpurun_out/r02_pybullet_probe.txt
PMG_COOP_STACK=1 timeout 900 python -m pytest tests/ -q -m gpu -s -k "block_stack or rearrange or variant or curriculum or decomposition or stack" > gpurun_out/gpu_tests_coop_stack.log 2>&1; tail -15 gpurun_out/gpu_tests_coop_stack.log
echo "== cooperative"; PMG_COOP_STACK=1 timeout 300 python tools/quick_time.py block_stack:2048 block_stack:256 block_stack:4096 2>&1 | grep -v "Task id"
echo "== thread-per-env"; timeout 300 python tools/quick_time.py block_stack:2048 block_stack:256 2>&1 | grep -v "Task id"
echo "== all GPU tests, default kernels"; timeout 1200 python -m pytest tests/ -q -m gpu > gpurun_out/gpu_tests_default.log 2>&1; tail -3 gpurun_out/gpu_tests_default.log
echo "== racecheck coop stack"; PMG_COOP_STACK=1 timeout 600 compute-sanitizer --tool racecheck python tools/race_stack.py 64 14 > gpurun_out/racecheck_coop_stack.log 2>&1; tail -4 gpurun_out/racecheck_coop_stack.log
