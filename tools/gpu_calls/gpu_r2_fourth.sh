# Round 2, fourth GPU call (2 GPUs): the sharded env's fused gather (2-GPU tests, bench fused vs NCCL, weak + strong),
# and timing / phase cycles of the multi-block kernel after hoisting the point classification out of the sweeps.
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 900 python -m pytest tests/test_gpu_sharded.py -q -m gpu -s > gpurun_out/gpu_tests_sharded.log 2>&1; tail -15 gpurun_out/gpu_tests_sharded.log | cut -c1-600
run2() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 "$@" 2>>gpurun_out/bench_2gpu.err | tail -1; }
run2 --steps 50 > gpurun_out/bench_reach_weak_2gpu_fused.json
run2 --steps 50 --gather nccl > gpurun_out/bench_reach_weak_2gpu_nccl.json
run2 --steps 50 --scaling strong > gpurun_out/bench_reach_strong_2gpu_fused.json
run2 --steps 50 --task block_stack --scaling strong > gpurun_out/bench_stack_strong_2gpu_fused.json
timeout 600 python bench.py --steps 50 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_reach_1gpu_b.json
python - <<'PY'
import json
for f in ("reach_1gpu_b", "reach_weak_2gpu_fused", "reach_weak_2gpu_nccl", "reach_strong_2gpu_fused", "stack_strong_2gpu_fused"):
    try:
        d = json.load(open("gpurun_out/bench_%s.json" % f))
        print(f, "value %.0f e2e %.0f ms/step %.3f kernel_ms %.3f launches %d breakdown %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["gpu_launches"], d["breakdown"]))
    except Exception as e:
        print(f, "FAILED", e)
PY
tail -5 gpurun_out/bench_2gpu.err
echo "== quick_time"; timeout 300 python tools/quick_time.py block_stack:2048 block_stack:256 block_stack:4096 block_rearrange:2048 2>&1 | grep -v "Task id"
export PMG_LIBRARY=pybullet_multigoal_gym_b200/libpmg_timing.so
for t in block_stack:2048; do echo "== $t"; timeout 300 python tools/coop_timing.py $t 2>&1 | grep -v "Task id"; done | tee gpurun_out/coop_timing_r2c.txt
unset PMG_LIBRARY
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_variants.py -q -m gpu -k "teacher or stack or rearrange" 2>&1 | tail -3
