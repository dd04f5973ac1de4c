#!/usr/bin/env python3
"""bench.py -- env-steps/sec (random policy) of the batched Kuka Reach env.step() hot path.

Contract: `python bench.py --gpus N --steps K --warmup W` (under torchrun for N > 1) prints ONE JSON
line on rank 0.  A "step" is one env.step() over the whole batch (action map + IK + 100 physics
substeps + observation/reward/flags for every environment), episodes of 50 steps with the reset
inside the timed region.  Workload: BASELINE.json configs[1], `reach` batch 8192 per GPU (weak
scaling: each rank owns 8192 environments; the only collective is the all-gather of the returned
observation batch).

  value        device-timed (CUDA events around every step, L2 flushed between steps outside the
               events), inputs resident in HBM, max over ranks
  e2e          the same metric through the public API with HOST buffers: env.step(numpy actions)
               -> pmg_step_host_blocks: H2D of the actions, the kernel, D2H of obs/reward/flags
  roofline     HBM roofline of the step kernel: algorithmic bytes per launch (SURVEY.md 8d, 282 B per
               Reach env-step) / mean kernel duration vs MEASURED_PEAKS.json hbm_gbs.  The path is
               bound by FP32 issue / dependency latency, not HBM; `fp32_frac` is
               reported beside it.
  cpu_baseline the CPU oracle (oracle/, a port: the reference's backend pybullet is not
               installable here) timed on all host cores on a bounded sample of the same workload.

`--impl reference` times that CPU port alone (there is no GPU work in that arm).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TASK = "reach"
BATCH_PER_GPU = 8192
EPISODE = 50
METRIC = "env-steps/sec (random policy) KukaReach batch=8192"
UNIT = "env-steps/s"
# algorithmic HBM bytes per env-step (fp32, 100 substeps fused in one kernel)
BYTES_PER_ENV_STEP = {"reach": 282, "push": 470, "pick_and_place": 474, "block_stack": 1138}
# order-of-magnitude useful FLOPs per env-step for the secondary FP32 fraction
FLOP_PER_ENV_STEP = {"reach": 1.5e6, "push": 3e6, "pick_and_place": 3e6, "block_stack": 7.5e6}
WORKLOAD = ("task=reach batch=8192 per GPU, 3-dim action (4th column ignored), state obs, sparse reward, "
            "50-step episodes, reset inside the timed region")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


def ncu_traffic(task):
    """dram bytes read+written per step-kernel launch from the committed ncu summary, or None."""
    p = os.path.join(ROOT, "profiles", "r01_step_kernel_ncu_summary.json")
    try:
        with open(p) as f:
            return json.load(f)[task]["dram_bytes_per_launch"]
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            pass

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "NVML unavailable"}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def cpu_port_rate(task, n_env, n_steps, threads):
    """env-steps/s of the CPU oracle (oracle/pmg_oracle.c) stepping n_env envs n_steps times."""
    import numpy as np
    from oracle import pmg_oracle as O
    O.build()
    envs = [O.OracleEnv(task, num_block=4, seed=i) for i in range(n_env)]
    for e in envs:
        e.reset()
    rng = np.random.RandomState(1234)
    actions = rng.uniform(-1, 1, size=(n_steps, n_env, envs[0].adim))
    secs = O.bench_rollout(envs, actions, threads)
    return n_env * n_steps / secs, secs


def run_reference(args):
    """Reference arm: the CPU port of the path on all host cores (no GPU work)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n_env = max(threads * 32, 64)  # bounded sample of the 8192-env batch: one pass = one "step"
    if args.warmup > 0:
        cpu_port_rate(TASK, n_env, min(args.warmup, 3), threads)
    rate, secs = cpu_port_rate(TASK, n_env, args.steps, threads)
    sample = "%d of %d envs x %d steps, %d pthreads, double precision C port (pybullet not installable)" % (
        n_env, BATCH_PER_GPU, args.steps, threads)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--task", default=TASK, help=argparse.SUPPRESS)
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU, help=argparse.SUPPRESS)
    ap.add_argument("--no-cpu-baseline", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.warmup < 3:
        args.warmup = 3

    import numpy as np
    import torch
    import torch.distributed as dist
    import pybullet_multigoal_gym_b200 as pmg
    from pybullet_multigoal_gym_b200.sharded import ShardedKukaEnv

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    distributed = world > 1
    if distributed:
        dist.init_process_group("nccl", device_id=dev)
    task, B, K, W = args.task, args.batch, args.steps, args.warmup

    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):  # make_env prints 'Task id: ...' like the reference
        if distributed:
            env_s = ShardedKukaEnv(task, B * world, device=local_rank, num_block=4, check_actions=False)
            env = env_s.env
        else:
            env_s = None
            env = pmg.make_env(task=task, batch=B, device=local_rank, num_block=4, check_actions=False)
    A, Wd = env.action_dim, env.row_width

    # synthetic random policy, resident in HBM before the timed region
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    tape = torch.rand((W + K, B, A), device=dev, generator=gen) * 2 - 1
    out = torch.empty((B, Wd), device=dev)
    reward = torch.empty((B,), device=dev)
    done = torch.empty((B,), dtype=torch.uint8, device=dev)
    success = torch.empty((B,), dtype=torch.uint8, device=dev)
    flush = torch.empty((256 << 20,), dtype=torch.uint8, device=dev)  # > 126 MB L2

    def one_step(t):
        if env_s is not None:
            env_s.step_gathered(tape[t])
        else:
            env.step_packed(tape[t], out, reward, done, success)

    elapsed_steps = 0

    def maybe_reset():
        nonlocal elapsed_steps
        elapsed_steps += 1
        if elapsed_steps % EPISODE == 0:
            env.reset(device_output=True)

    env.reset(device_output=True)
    for t in range(W):
        one_step(t)
        maybe_reset()
    torch.cuda.synchronize()
    if distributed:
        dist.barrier()
    launches0 = env.launch_count
    sampler = ClockSampler(local_rank)
    sampler.start()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    kstops = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    torch.cuda.synchronize()
    for k in range(K):
        flush.zero_()  # L2 flush, outside the timed events
        starts[k].record()
        one_step(W + k)
        kstops[k].record()  # end of the step kernel (+ all-gather when sharded)
        maybe_reset()       # the episode reset is part of the rollout, inside the timed region
        stops[k].record()
    torch.cuda.synchronize()
    if distributed:
        dist.barrier()
    sampler.stop_flag = True
    sampler.join()
    total_ms = sum(s.elapsed_time(e) for s, e in zip(starts, stops))
    kernel_ms = sum(s.elapsed_time(e) for s, e in zip(starts, kstops)) / K
    launches = env.launch_count - launches0
    overflow = env.overflow_count
    t_ms = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if distributed:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    total_ms_max = float(t_ms.item())
    value = B * world * K / (total_ms_max / 1e3)

    # ---- e2e: public API, host buffers, H2D + kernel + D2H every step ------------------------
    host_tape = tape[W:W + min(K, 50)].cpu().numpy()
    env.reset(device_output=True)
    env.step(host_tape[0])
    torch.cuda.synchronize()
    if distributed:
        dist.barrier()
    t0 = time.perf_counter()
    for k in range(host_tape.shape[0]):
        env.step(host_tape[k])
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e_t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if distributed:
        dist.all_reduce(e_t, op=dist.ReduceOp.MAX)
    e2e_value = B * world * host_tape.shape[0] / float(e_t.item())

    if rank == 0:
        hbm_peak, peak_src, sm_max_mhz = peaks()
        achieved = BYTES_PER_ENV_STEP[task] * B / (kernel_ms / 1e3) / 1e9
        fp32_peak = 148 * 128 * 2 * sm_max_mhz * 1e6
        line = {
            "metric": METRIC if (task == TASK and B == BATCH_PER_GPU) else "env-steps/sec (random policy) %s batch=%d" % (task, B),
            "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": total_ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD if task == TASK else "task=%s batch=%d per GPU" % (task, B),
                       "global_batch": B * world, "parallelism": "env-sharded x%d, one all-gather of the obs batch per step" % world,
                       "l2": "flushed between timed steps (256 MiB memset outside the per-step CUDA events); working set 1.6 MB",
                       "kernel": ("lane-cooperative step kernel: 8 lanes per env, 4 envs per one-warp block"
                                  if ((task == "reach" and os.environ.get("PMG_COOP", "1") != "0") or
                                      (task in ("push", "pick_and_place") and os.environ.get("PMG_COOP_BLOCK", "1") != "0"))
                                  else "thread-per-env step kernel: 32 envs per warp"),
                       "contact_pool_overflows": overflow},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "traffic": ncu_traffic(task), "peak_source": peak_src,
                         "bytes_per_env_step": BYTES_PER_ENV_STEP[task], "kernel_ms": kernel_ms,
                         "note": "latency/FP32-issue bound path: 100 dependent substeps per env-step, ~0.3 KB of compulsory HBM traffic (SURVEY.md 0.5)",
                         "fp32_frac": FLOP_PER_ENV_STEP[task] * B / (kernel_ms / 1e3) / fp32_peak},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * A * 4, "d2h_bytes_per_step": B * Wd * 4 + B * 4 + 2 * B,
                    "steps": int(host_tape.shape[0]), "api": "env.step(numpy) -> pmg_step_host_blocks"},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
        }
        if not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            n_env, n_steps = max(threads * 32, 64), 100
            rate, secs = cpu_port_rate(task, n_env, n_steps, threads)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": "%d of %d envs x %d steps (%.1f s), %d pthreads, double-precision C port of the path; the reference's pybullet backend is not installable here" % (n_env, B, n_steps, secs, threads)}
        print(json.dumps(line), flush=True)
    if distributed:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
