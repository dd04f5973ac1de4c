mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -q -m gpu -s > gpurun_out/parity.log 2>&1
tail -5 gpurun_out/parity.log; grep -E "worst|error:" gpurun_out/parity.log
python tools/quick_time.py 2>&1 | grep -v "Task id" | tee gpurun_out/quick_time.log
