# Round 2, fifth GPU call (1 GPU): environments-per-block selection + register-resident point classification.
mkdir -p gpurun_out
echo "== quick_time (default geometry)"; timeout 600 python tools/quick_time.py block_stack:2048 block_stack:256 block_stack:4096 block_rearrange:2048 reach:8192 reach:1024 push:4096 push:512 pick_and_place:4096 pick_and_place:512 2>&1 | grep -v "Task id"
echo "== quick_time PMG_COOP_EPB=4"; PMG_COOP_EPB=4 timeout 600 python tools/quick_time.py block_stack:2048 block_stack:256 reach:1024 pick_and_place:512 2>&1 | grep -v "Task id"
echo "== quick_time PMG_COOP_EPB=1"; PMG_COOP_EPB=1 timeout 600 python tools/quick_time.py block_stack:2048 block_stack:256 2>&1 | grep -v "Task id"
export PMG_LIBRARY=pybullet_multigoal_gym_b200/libpmg_timing.so
for t in block_stack:2048; do echo "== $t"; PMG_COOP_EPB=4 timeout 300 python tools/coop_timing.py $t 2>&1 | grep -v "Task id"; done | tee gpurun_out/coop_timing_r2d.txt
unset PMG_LIBRARY
timeout 900 python -m pytest tests/ -q -m gpu 2>&1 | tail -3
timeout 900 python bench.py --task block_stack --steps 50 2>gpurun_out/bench_block_stack.err | tail -1 > gpurun_out/bench_block_stack.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_block_stack.json"))
print("block_stack value %.0f e2e %.0f ms/step %.3f kernel_ms %.3f cpu %s overflow %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d.get("cpu_baseline", {}).get("value"), d["config"]["contact_pool_overflows"]))
PY
