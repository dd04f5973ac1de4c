mkdir -p gpurun_out
echo "== tiled (default)"; python tools/reward_bw.py 2>&1 | tee gpurun_out/reward_bw_tiled.txt
echo "== row per thread (PMG_REWARD_SIMPLE=1)"; PMG_REWARD_SIMPLE=1 python tools/reward_bw.py 2>&1 | tee gpurun_out/reward_bw_simple.txt
python -m pytest tests/ -q -m gpu -k "reward or her" 2>&1 | tail -2
