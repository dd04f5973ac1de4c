# Final validation of the round: tests, smoke, both bench arms, launch list, ncu captures (CSV exports made on the
# box), soak, phase timing, sanitizers.
mkdir -p gpurun_out
python -m pytest tests/ -q -m gpu -s > gpurun_out/gpu_tests.log 2>&1; tail -2 gpurun_out/gpu_tests.log; grep -E "teacher-forced|worst|golden rollout|resting|cooperative vs" gpurun_out/gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/clocks.csv &
SMI=$!
python bench.py 2>&1 | tail -1 > gpurun_out/bench_default.json
kill $SMI
python bench.py --impl reference 2>&1 | tail -1 > gpurun_out/bench_reference.json
python -c "
import json
d=json.load(open('gpurun_out/bench_default.json')); r=json.load(open('gpurun_out/bench_reference.json'))
print('value %.0f e2e %.0f ms/step %.3f kernel_ms %.3f launches %d clocks %s cpu %.0f (%d cores) ref-arm %.0f roofline frac %.2e fp32 %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['gpu_launches'], d['clocks'], d['cpu_baseline']['value'], d['cpu_baseline']['cores'], r['value'], d['roofline']['frac'], d['roofline']['fp32_frac']))"
for t in push pick_and_place; do
  python bench.py --task $t --batch 4096 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_$t.json
  PMG_COOP_BLOCK=0 python bench.py --task $t --batch 4096 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_${t}_thread.json
  python -c "
import json
for f in ('gpurun_out/bench_$t.json', 'gpurun_out/bench_${t}_thread.json'):
    d=json.load(open(f)); print('$t', d['config']['kernel'][:20], 'value %.0f e2e %.0f ms/step %.3f overflow %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['config']['contact_pool_overflows']))"
done
python bench.py --task block_stack --batch 2048 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_block_stack.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 2 -c 1 -o /tmp/prof_coop_reach -f python tools/prof_one.py reach 8192 4 > gpurun_out/ncu_coop_reach.log 2>&1; tail -1 gpurun_out/ncu_coop_reach.log
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 11 -c 1 -o /tmp/prof_coop_reach_down -f python tools/prof_one.py reach 8192 12 down > gpurun_out/ncu_coop_reach_down.log 2>&1; tail -1 gpurun_out/ncu_coop_reach_down.log
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 6 -c 1 -o /tmp/prof_coop_push4096 -f python tools/prof_one.py push 4096 8 > gpurun_out/ncu_coop_push4096.log 2>&1; tail -1 gpurun_out/ncu_coop_push4096.log
for n in coop_reach coop_reach_down coop_push4096; do
  ncu -i /tmp/prof_$n.ncu-rep --page raw --csv > gpurun_out/prof_${n}_raw.csv
  ncu -i /tmp/prof_$n.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/prof_${n}_source.csv
  cp /tmp/prof_$n.ncu-rep gpurun_out/
done
python tools/soak.py 2048 3 2>&1 | grep -v "Task id" | tee gpurun_out/soak.txt
timeout 200 compute-sanitizer --tool memcheck python tools/prof_one.py block_rearrange 64 3 > gpurun_out/memcheck_rearrange.log 2>&1; tail -1 gpurun_out/memcheck_rearrange.log
timeout 300 compute-sanitizer --tool racecheck python tools/prof_one.py reach 16 11 down > gpurun_out/racecheck_coop.log 2>&1; tail -1 gpurun_out/racecheck_coop.log
timeout 300 compute-sanitizer --tool racecheck python tools/prof_one.py pick_and_place 16 6 > gpurun_out/racecheck_coop_pnp.log 2>&1; tail -1 gpurun_out/racecheck_coop_pnp.log
timeout 300 compute-sanitizer --tool memcheck python tools/prof_one.py push 64 6 > gpurun_out/memcheck_coop_push.log 2>&1; tail -1 gpurun_out/memcheck_coop_push.log
bash tools/gpu_calls/gpu_timing.sh | tee gpurun_out/coop_timing.txt
bash tools/gpu_calls/gpu_timing_blk.sh | tee gpurun_out/coop_timing_blk.txt
