# Round 2, final 1-GPU measurement call: GPU tests with their printed parity statistics, smoke, ncu instruction counts of
# the step kernels in steady state (-> profiles/r02_step_kernel_ncu_summary.json, which bench.py's issue_slot_frac reads),
# bench lines of every task, the reference arm, the ncu launch list of the default bench command.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -q -m gpu -s > gpurun_out/gpu_tests.log 2>&1; tail -3 gpurun_out/gpu_tests.log; grep -E "teacher-forced|golden rollout|resting|cooperative vs|FAILED|Error" gpurun_out/gpu_tests.log | cut -c1-500
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
M=smsp__inst_executed.sum,smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sass__inst_executed_local_loads,sass__inst_executed_local_stores,sm__icc_request_hit_rate.pct,sm__warps_active.avg.per_cycle_active
ARGS=""
for tb in reach:8192 push:4096 pick_and_place:4096 slide:4096 block_stack:2048; do
  t=${tb%%:*}; b=${tb##*:}
  timeout 600 ncu --metrics $M --clock-control none -k regex:step_kernel -s 62 -c 3 --csv --log-file gpurun_out/ncu_counts_${t}_$b.csv python tools/prof_steady.py $t $b 5 > /dev/null 2>&1
  ARGS="$ARGS ${t}_$b=gpurun_out/ncu_counts_${t}_$b.csv"
done
python tools/ncu_counts.py profiles/r02_step_kernel_ncu_summary.json $ARGS > gpurun_out/ncu_counts_summary.txt 2>&1; cp profiles/r02_step_kernel_ncu_summary.json gpurun_out/
for k in 20 100; do timeout 600 python bench.py --steps $k --warmup 5 2>gpurun_out/bench_reach_$k.err | tail -1 > gpurun_out/bench_reach_$k.json; done
timeout 400 python bench.py --impl reference --steps 20 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_reference_20.json
for t in push pick_and_place slide block_stack; do timeout 900 python bench.py --task $t --steps 50 2>gpurun_out/bench_$t.err | tail -1 > gpurun_out/bench_$t.json; done
python - <<'PY'
import json
for f in ("reach_20", "reach_100", "push", "pick_and_place", "slide", "block_stack"):
    try:
        d = json.load(open("gpurun_out/bench_%s.json" % f))
        print(f, "value %.0f e2e %.0f ms/step %.3f kernel_ms %.3f launches %d cpu %s overflow %s issue %.3f clocks %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["gpu_launches"], d.get("cpu_baseline", {}).get("value"), d["config"]["contact_pool_overflows"], d["roofline"].get("issue_slot_frac", -1), d["clocks"]))
    except Exception as e:
        print(f, "FAILED", e)
print(open("gpurun_out/bench_reference_20.json").read()[:300])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 compute-sanitizer --tool racecheck python tools/race_stack.py 64 14 > gpurun_out/racecheck_coop_stack.log 2>&1; tail -3 gpurun_out/racecheck_coop_stack.log
