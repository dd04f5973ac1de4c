# Round 2, last 1-GPU call on the final tree: GPU tests, smoke, the driver's two bench commands (both arms), the 100-step
# line, the ncu launch list of the default bench command.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -q -m gpu -s > gpurun_out/gpu_tests_30.log 2>&1; tail -2 gpurun_out/gpu_tests_30.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 2>/dev/null | tail -1 > gpurun_out/bench30_reference_arm.json
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 2>gpurun_out/bench30.err | tail -1 > gpurun_out/bench30_reach_20.json
timeout 400 python bench.py --steps 100 --warmup 5 --no-cpu-baseline 2>>gpurun_out/bench30.err | tail -1 > gpurun_out/bench30_reach_100.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches30.csv python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench30_under_ncu.log 2>&1
python - <<'PY'
import json
for f in ("reach_20", "reach_100"):
    d = json.load(open("gpurun_out/bench30_%s.json" % f))
    print(f, "value %.0f e2e %.0f ms/step %.3f launches %d cpu %s clocks %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["gpu_launches"], d.get("cpu_baseline", {}).get("value"), d["clocks"]))
print(open("gpurun_out/bench30_reference_arm.json").read()[:400])
PY
