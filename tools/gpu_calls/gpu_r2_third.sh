# Round 2, third GPU call (1 GPU): re-run the GPU tests after the test / sweep rework, debug the grasp velocity entries,
# phase cycles + timing of the reworked cooperative multi-block kernel, bench line for block_stack.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -q -m gpu -s > gpurun_out/gpu_tests.log 2>&1; tail -3 gpurun_out/gpu_tests.log; grep -E "teacher-forced|FAILED|Error" gpurun_out/gpu_tests.log | cut -c1-700
timeout 600 python tools/debug_grasp_vel.py 2>&1 | grep -v "Task id" | tail -60 > gpurun_out/debug_grasp_vel.txt; cat gpurun_out/debug_grasp_vel.txt | head -50
echo "== quick_time"; timeout 300 python tools/quick_time.py block_stack:2048 block_stack:256 block_stack:4096 block_rearrange:2048 2>&1 | grep -v "Task id"
timeout 900 python bench.py --task block_stack --steps 50 2>gpurun_out/bench_block_stack.err | tail -1 > gpurun_out/bench_block_stack.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_block_stack.json"))
print("block_stack value %.0f e2e %.0f ms/step %.3f kernel_ms %.3f cpu %s overflow %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d.get("cpu_baseline", {}).get("value"), d["config"]["contact_pool_overflows"]))
PY
export PMG_LIBRARY=pybullet_multigoal_gym_b200/libpmg_timing.so
for t in block_stack:256 block_stack:2048; do echo "== $t"; timeout 300 python tools/coop_timing.py $t 2>&1 | grep -v "Task id"; done | tee gpurun_out/coop_timing_r2b.txt
unset PMG_LIBRARY
timeout 600 compute-sanitizer --tool racecheck python tools/race_stack.py 64 14 > gpurun_out/racecheck_coop_stack2.log 2>&1; tail -3 gpurun_out/racecheck_coop_stack2.log
timeout 600 compute-sanitizer --tool memcheck python tools/race_stack.py 64 14 > gpurun_out/memcheck_coop_stack2.log 2>&1; tail -3 gpurun_out/memcheck_coop_stack2.log
