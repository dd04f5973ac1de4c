# Re-run of the file-producing parts of gpu_final4.sh with small outputs (the .ncu-rep files stay on the box: only
# their CSV exports come back).
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/clocks.csv &
SMI=$!
python bench.py 2>&1 | tail -1 > gpurun_out/bench_default.json
kill $SMI
python bench.py --impl reference 2>&1 | tail -1 > gpurun_out/bench_reference.json
for t in push pick_and_place; do
  python bench.py --task $t --batch 4096 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_$t.json
  PMG_COOP_BLOCK=0 python bench.py --task $t --batch 4096 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_${t}_thread.json
done
python bench.py --task block_stack --batch 2048 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_block_stack.json
python -c "
import json
for n in ('default','reference','push','push_thread','pick_and_place','pick_and_place_thread','block_stack'):
    d=json.load(open('gpurun_out/bench_%s.json' % n)); print(n, 'value %.0f e2e %.0f ms/step %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']))"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 2 -c 1 -o /tmp/prof_coop_reach -f python tools/prof_one.py reach 8192 4 > gpurun_out/ncu_coop_reach.log 2>&1; tail -1 gpurun_out/ncu_coop_reach.log
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 11 -c 1 -o /tmp/prof_coop_reach_down -f python tools/prof_one.py reach 8192 12 down > gpurun_out/ncu_coop_reach_down.log 2>&1; tail -1 gpurun_out/ncu_coop_reach_down.log
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 6 -c 1 -o /tmp/prof_coop_push4096 -f python tools/prof_one.py push 4096 8 > gpurun_out/ncu_coop_push4096.log 2>&1; tail -1 gpurun_out/ncu_coop_push4096.log
for n in coop_reach coop_reach_down coop_push4096; do
  ncu -i /tmp/prof_$n.ncu-rep --page raw --csv > gpurun_out/prof_${n}_raw.csv
  ncu -i /tmp/prof_$n.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/prof_${n}_source.csv
done
du -sh gpurun_out
