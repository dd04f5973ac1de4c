mkdir -p gpurun_out
python -m pytest tests/ -q -m gpu -s > gpurun_out/gpu_tests.log 2>&1; tail -2 gpurun_out/gpu_tests.log; grep -E "joint-control|_jc|Error|assert " gpurun_out/gpu_tests.log | head
python bench.py --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('reach default value %.0f e2e %.0f ms %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']))"
echo "== coop jc"; python tools/quick_time.py reach:8192:jc pick_and_place:4096:jc push:4096:jc 2>&1 | grep -v "Task id"
echo "== thread jc"; PMG_COOP=0 PMG_COOP_BLOCK=0 python tools/quick_time.py reach:8192:jc pick_and_place:4096:jc push:4096:jc 2>&1 | grep -v "Task id"
