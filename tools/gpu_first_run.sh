mkdir -p gpurun_out
python -m pytest tests/ -q -m gpu -s > gpurun_out/parity.log 2>&1
tail -5 gpurun_out/parity.log; grep -E "worst|teacher-forced:" gpurun_out/parity.log; grep -E "^E   .*(Error|assert)" gpurun_out/parity.log | head -10
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 20 --warmup 3 2>&1 | tail -2 | tee gpurun_out/bench_quick.json
python bench.py --impl reference --steps 20 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_ref_quick.json
