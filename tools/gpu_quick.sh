python -m pytest tests/ -q -m gpu 2>&1 | tail -2
python bench.py --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('BENCH value %.3f M/s  ms/step %.3f  kernel_ms %.3f  e2e %.3f M/s' % (d['value']/1e6, d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['value']/1e6))"
