mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 50 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_8gpu.json | cut -c1-400
