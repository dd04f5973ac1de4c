# Round 2, sixth GPU call (1 GPU): Slide on hardware (parity suite, timing, bench line), steady-state phase cycles of Reach.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/ -q -m gpu 2>&1 | tail -25 | tee gpurun_out/r2_sixth_tests.txt
echo "== quick_time"; timeout 600 python tools/quick_time.py slide:4096 slide:512 push:4096 reach:8192 2>&1 | grep -v "Task id"
timeout 900 python bench.py --task slide --steps 50 2>gpurun_out/bench_slide.err | tail -1 > gpurun_out/bench_slide.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_slide.json"))
print("slide value %.0f e2e %.0f ms/step %.3f kernel_ms %.3f cpu %s overflow %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d.get("cpu_baseline", {}).get("value"), d["config"]["contact_pool_overflows"]))
PY
export PMG_LIBRARY=pybullet_multigoal_gym_b200/libpmg_timing.so
for t in reach:8192 slide:4096; do echo "== $t"; timeout 300 python tools/coop_timing.py $t 2>&1 | grep -v "Task id"; done | tee gpurun_out/coop_timing_r2e.txt
unset PMG_LIBRARY
