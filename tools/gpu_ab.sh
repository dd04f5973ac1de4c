run() { python bench.py --steps 50 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 value %.3f M/s ms/step %.3f' % (d['value']/1e6, d['ms_per_step']))"; }
PMG_BULK_COPY=1 run "bulk+maxL1      "
run "nobulk+maxL1    "
PMG_BULK_COPY=1 PMG_DEFAULT_CARVEOUT=1 run "bulk+default    "
PMG_DEFAULT_CARVEOUT=1 run "nobulk+default  "
python tools/quick_time.py 2>&1 | grep -v "Task id"
