# 2-GPU call: sharded-env tests (gathered rows == single-GPU rows) and a short bench line after the peer-connect rework
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sharded.py -q -m gpu -x 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29733 bench.py --gpus 2 --steps 20 --warmup 5 2>gpurun_out/bench29_2gpu.err | tail -1 > gpurun_out/bench29_reach_weak_2gpu.json
python -c "
import json; d=json.load(open('gpurun_out/bench29_reach_weak_2gpu.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['config']['parallelism'])"
tail -3 gpurun_out/bench29_2gpu.err
