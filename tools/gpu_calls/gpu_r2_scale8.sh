# Round 2, 8-GPU call: weak scaling of the headline config (fused gather vs NCCL all-gather) and the SHARDED configs of
# BASELINE.json (strong scaling: reach 8192 -> 1024/GPU, push / pick_and_place 4096 -> 512/GPU, block_stack 2048 -> 256/GPU).
mkdir -p gpurun_out
nvidia-smi -L | wc -l
run8() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 8 "$@" 2>>gpurun_out/bench_8gpu.err | tail -1; }
run8 --steps 100 > gpurun_out/bench_reach_weak_8gpu_fused.json
run8 --steps 100 --gather nccl > gpurun_out/bench_reach_weak_8gpu_nccl.json
run8 --steps 50 --scaling strong > gpurun_out/bench_reach_strong_8gpu_fused.json
run8 --steps 50 --task pick_and_place --scaling strong > gpurun_out/bench_pick_and_place_strong_8gpu_fused.json
run8 --steps 50 --task push --scaling strong > gpurun_out/bench_push_strong_8gpu_fused.json
run8 --steps 50 --task block_stack --scaling strong > gpurun_out/bench_block_stack_strong_8gpu_fused.json
python - <<'PY'
import json
for f in ("reach_weak_8gpu_fused", "reach_weak_8gpu_nccl", "reach_strong_8gpu_fused", "pick_and_place_strong_8gpu_fused", "push_strong_8gpu_fused", "block_stack_strong_8gpu_fused"):
    try:
        d = json.load(open("gpurun_out/bench_%s.json" % f))
        b = d["breakdown"]
        print(f, "value %.0f e2e %.0f ms/step %.3f launches %d kernel_ms max %.3f min %.3f total_ms max %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["gpu_launches"],
              max(b["step_kernel_ms_per_rank"]), min(b["step_kernel_ms_per_rank"]), max(b["step_total_ms_per_rank"])))
    except Exception as e:
        print(f, "FAILED", e)
PY
tail -5 gpurun_out/bench_8gpu.err
