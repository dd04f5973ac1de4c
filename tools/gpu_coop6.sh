mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -q -m gpu -k "reach" -s 2>&1 | grep -E "worst|passed|failed" 
echo "== timeline coop"; python tools/step_timeline.py reach 8192 2>&1 | grep -v "Task id"
python tools/quick_time.py reach:8192 reach:65536 2>&1 | grep -v "Task id"
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 11 -c 1 -o gpurun_out/prof_coop_reach_down -f python tools/prof_one.py reach 8192 12 down > gpurun_out/ncu_coop_reach_down.log 2>&1; tail -1 gpurun_out/ncu_coop_reach_down.log
