# Round 2, call 26: GPU tests on the build with the block_rearrange curriculum; default bench line; throughput at batches
# beyond the configs' (where the step is issue-bound instead of bound by its slowest environment).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -q -m gpu -s > gpurun_out/gpu_tests_26.log 2>&1; tail -3 gpurun_out/gpu_tests_26.log; grep -E "FAILED|Error|curriculum" gpurun_out/gpu_tests_26.log | cut -c1-300 | head
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 python bench.py --steps 20 --warmup 5 2>gpurun_out/bench26_reach.err | tail -1 > gpurun_out/bench26_reach_8192.json
for tb in reach:32768 reach:65536 push:16384 pick_and_place:16384 block_stack:8192; do
  t=${tb%%:*}; b=${tb##*:}
  timeout 300 python bench.py --task $t --batch $b --steps 30 --warmup 5 --no-cpu-baseline 2>gpurun_out/bench26_${t}_$b.err | tail -1 > gpurun_out/bench26_${t}_$b.json
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench26_*.json")):
    try:
        d = json.load(open(f))
        print(f.split("bench26_")[1], "value %.0f e2e %.0f ms/step %.3f issue %.3f overflow %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"].get("issue_slot_frac", -1), d["config"]["contact_pool_overflows"]))
    except Exception as e:
        print(f, "FAILED", e)
PY
