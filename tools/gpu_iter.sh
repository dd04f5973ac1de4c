mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "reach_rollout or teacher or golden" -s > gpurun_out/parity.log 2>&1
tail -3 gpurun_out/parity.log; grep -E "worst|teacher-forced:" gpurun_out/parity.log
python bench.py --steps 50 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('BENCH value %.3f M/s  ms/step %.3f  kernel_ms %.3f  e2e %.3f M/s' % (d['value']/1e6, d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['value']/1e6))"
python tools/quick_time.py 2>&1 | grep -v "Task id"
