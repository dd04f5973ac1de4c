# Round 2, second GPU call (1 GPU): all GPU tests, bench lines for configs 2-5 in steady state, --steps 20 vs 200 check,
# ncu instruction counts + full captures at steady state, phase cycles of the cooperative multi-block kernel.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -q -m gpu -s > gpurun_out/gpu_tests.log 2>&1; tail -3 gpurun_out/gpu_tests.log; grep -E "teacher-forced|golden rollout|resting|cooperative vs|FAILED|Error" gpurun_out/gpu_tests.log | cut -c1-400
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for k in 20 200; do timeout 600 python bench.py --steps $k --warmup 5 2>gpurun_out/bench_reach_$k.err | tail -1 > gpurun_out/bench_reach_$k.json; done
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_reference_20.json
for t in push pick_and_place block_stack; do timeout 900 python bench.py --task $t --steps 50 2>gpurun_out/bench_$t.err | tail -1 > gpurun_out/bench_$t.json; done
python - <<'PY'
import json
for f in ("reach_20", "reach_200", "push", "pick_and_place", "block_stack"):
    try:
        d = json.load(open("gpurun_out/bench_%s.json" % f))
        print(f, "value %.0f e2e %.0f ms/step %.3f kernel_ms %.3f launches %d cpu %s overflow %s clocks %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["gpu_launches"], d.get("cpu_baseline", {}).get("value"), d["config"]["contact_pool_overflows"], d["clocks"]))
    except Exception as e:
        print(f, "FAILED", e)
print(open("gpurun_out/bench_reference_20.json").read()[:300])
PY
M=smsp__inst_executed.sum,smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sass__inst_executed_local_loads,sass__inst_executed_local_stores,sm__icc_request_hit_rate.pct,sm__warps_active.avg.per_cycle_active
for tb in reach:8192 push:4096 pick_and_place:4096 block_stack:2048; do
  t=${tb%%:*}; b=${tb##*:}
  timeout 600 ncu --metrics $M --clock-control none -k regex:step_kernel -s 62 -c 3 --csv --log-file gpurun_out/ncu_counts_${t}_$b.csv python tools/prof_steady.py $t $b 5 > /dev/null 2>&1
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
for tb in reach:8192 block_stack:2048; do
  t=${tb%%:*}; b=${tb##*:}
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 62 -c 1 -o /tmp/prof_$t -f python tools/prof_steady.py $t $b 3 > gpurun_out/ncu_full_$t.log 2>&1; tail -1 gpurun_out/ncu_full_$t.log
  ncu -i /tmp/prof_$t.ncu-rep --page raw --csv > gpurun_out/prof_${t}_raw.csv
  ncu -i /tmp/prof_$t.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/prof_${t}_source.csv
done
export PMG_LIBRARY=pybullet_multigoal_gym_b200/libpmg_timing.so
for t in block_stack:256 block_stack:2048 reach:8192; do echo "== $t"; timeout 300 python tools/coop_timing.py $t 2>&1 | grep -v "Task id"; done | tee gpurun_out/coop_timing_r2.txt
