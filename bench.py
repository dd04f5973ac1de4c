#!/usr/bin/env python3
"""bench.py -- env-steps/sec (random policy) of the batched Kuka env.step() hot path.

Contract: `python bench.py --gpus N --steps K --warmup W` (under torchrun for N > 1) prints ONE JSON line on rank 0.
A "step" is one env.step() over the whole batch: action map + IK + 100 physics substeps + observation / reward /
flags for every environment, and the reset of the environments whose 50-step episode ended in that step.

Workload (default = BASELINE.json configs[1]): `reach`, batch 8192 per GPU (weak scaling).  `--task push|
pick_and_place|block_stack|block_rearrange --batch B` select configs[2..4]; `--scaling strong` shards the batch named by
--batch over the N ranks (configs[3], [4]: 4096 -> 512 per GPU, 2048 -> 256 per GPU) instead of giving every rank
--batch environments.

STEADY STATE for any --steps: the episodes are staggered (environment i starts at elapsed step i mod 50), the
environments reset themselves on the device when their episode ends (Philox-sampled spawn rows, pmg_set_auto_reset:
B/50 resets inside EVERY timed step, in the timed events), and 60 untimed set-up steps precede the warm-up so that
the contact state of the batch (arms resting on the table, blocks pushed around) has reached its stationary mix.

  value        device-timed: CUDA events around every step (step kernel + auto-reset pass [+ the gather over peer
               memory when sharded]), L2 flushed between steps outside the events, inputs resident in HBM, max over ranks
  e2e          the same metric through the public API with HOST buffers: env.step(numpy) -> pmg_step_host_blocks
               (N = 1) / ShardedKukaEnv.step_host (N > 1: H2D of the local actions, step + gather, D2H of the GLOBAL batch)
  roofline     HBM roofline of the step kernel: algorithmic bytes per launch / its mean duration, timed
               by CUDA events the library records around that kernel alone on its stream, vs MEASURED_PEAKS.json.
               The path is bound by warp-instruction issue / dependent-issue latency, not HBM:
               `issue_slot_frac` and `fp32_frac` are derived from ncu-MEASURED instruction counts of the same kernel at
               the same batch (profiles/r02_step_kernel_ncu_summary.json) and the live kernel time.
  cpu_baseline the CPU oracle (oracle/, a double-precision C port: the reference's backend pybullet is not installable
               here or on the GPU box -- profiles/r02_pybullet_probe.txt) on all host cores, same staggered workload,
               compiled -O3 -march=native on the box, warmed up, on a bounded sample of the batch.

`--impl reference` times that CPU port alone (no GPU work in that arm).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

EPISODE = 50
SETTLE_STEPS = 60
UNIT = "env-steps/s"
DEFAULT_BATCH = {"reach": 8192, "push": 4096, "pick_and_place": 4096, "block_stack": 2048, "block_rearrange": 2048, "slide": 4096}
HEADLINE = "env-steps/sec (random policy) KukaReach batch=8192"
# algorithmic HBM bytes per env-step (fp32, 100 substeps fused in one kernel)
BYTES_PER_ENV_STEP = {"reach": 282, "push": 470, "pick_and_place": 474, "block_stack": 1138, "block_rearrange": 1134, "slide": 470}


def workload_text(task, per_gpu, world, scaling):
    return ("task=%s batch=%d per GPU x %d GPU(s) (%s scaling), random policy, state obs, 50-step episodes staggered over the "
            "batch (env i starts at step i mod 50), device-side auto-reset inside every timed step, %d untimed set-up steps "
            "before the warm-up" % (task, per_gpu, world, scaling, SETTLE_STEPS))


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_summary(task, per_gpu):
    """ncu-measured per-launch counters of the step kernel at this batch (committed under profiles/), or {}."""
    for name in ("r02_step_kernel_ncu_summary.json", "r01_step_kernel_ncu_summary.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                d = json.load(f)
            for key in ("%s_%d" % (task, per_gpu), task):
                if key in d and (key != task or d[key].get("batch", per_gpu) == per_gpu):
                    out = dict(d[key])
                    out["source"] = "profiles/" + name
                    return out
        except Exception:
            pass
    return {}


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
            time.sleep(0.02)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "NVML unavailable"}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def cpu_port_rate(task, n_env, n_steps, threads, warm_steps):
    """env-steps/s of the CPU oracle stepping n_env staggered envs n_steps times after warm_steps untimed steps
    (its rollout driver resets an env when its episode ends, like the auto-reset of the GPU arm)."""
    import numpy as np
    from oracle import pmg_oracle as O
    flags = O.use_native()
    envs = [O.OracleEnv(task, num_block=4, seed=i) for i in range(n_env)]
    for i, e in enumerate(envs):
        e.reset()
        st = e.get_state()
        st[-1] = i % EPISODE
        e.set_state(st)
    rng = np.random.RandomState(1234)
    adim = envs[0].adim
    if warm_steps > 0:
        O.bench_rollout(envs, rng.uniform(-1, 1, size=(warm_steps, n_env, adim)), threads)
    secs = O.bench_rollout(envs, rng.uniform(-1, 1, size=(n_steps, n_env, adim)), threads)
    return n_env * n_steps / secs, secs, flags


def cpu_sample_size(task, threads, steps=None):
    """Environments of the CPU sample.  With `steps` (the reference arm: the driver passes few) the sample grows until the
    timed region holds ~2 s of work on 16 cores, so that thread start-up and the last thread's tail do not weigh on the
    rate (a 0.4 s region read 20 % low)."""
    per_thread = {"reach": 32, "push": 16, "pick_and_place": 16, "slide": 16}.get(task, 8)
    n = max(threads * per_thread, 64)
    if steps:
        target = {"reach": 60000, "push": 40000, "pick_and_place": 40000, "slide": 40000}.get(task, 24000)  # env-steps in the timed region
        want = -(-target // max(steps, 1))
        want = -(-want // threads) * threads
        n = max(n, min(want, 4096))
    return n


def run_reference(args, task, per_gpu):
    """Reference arm: the CPU port of the path on all host cores (no GPU work).  One "step" = one env.step() of every
    environment of a bounded sample of the batch; the rate is what the metric counts (env-steps/s)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n_env = min(cpu_sample_size(task, threads, max(args.steps, 1)), per_gpu)
    warm = max(args.warmup, 60)  # settle the contact mix and the CPU clocks / caches before timing
    steps = max(args.steps, 1)
    rate, secs, flags = cpu_port_rate(task, n_env, steps, threads, warm)
    sample = ("%d of %d envs x %d steps after %d untimed warm-up steps, %d pthreads, double-precision C port built %s; the "
              "reference's own backend (pybullet) is not installable here or on the GPU box" % (n_env, per_gpu, steps, warm, threads, flags))
    line = {
        "impl": "reference", "metric": metric_name(task, per_gpu), "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * secs / steps, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_text(task, per_gpu, 1, args.scaling), "sample": sample},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def metric_name(task, per_gpu):
    if task == "reach" and per_gpu == 8192:
        return HEADLINE
    return "env-steps/sec (random policy) %s batch=%d per GPU" % (task, per_gpu)


E2E_STEPS = 50  # steps of the end-to-end leg, whatever --steps is


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--task", default="reach", choices=sorted(DEFAULT_BATCH), help="BASELINE.json configs[1..4] (+ block_rearrange)")
    ap.add_argument("--batch", type=int, default=None, help="environments per GPU (weak) or in total (strong); default: the config's batch")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: every rank steps --batch environments; strong: --batch is sharded over the ranks")
    ap.add_argument("--gather", default="fused", choices=["fused", "nccl"], help="N > 1: peer-memory gather in the step (default) or one NCCL all-gather")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    task = args.task
    world_env = int(os.environ.get("WORLD_SIZE", "1"))
    batch = args.batch if args.batch is not None else DEFAULT_BATCH[task]
    if args.scaling == "strong":
        if batch % max(world_env, 1):
            raise SystemExit("--scaling strong: --batch %d is not divisible by the %d ranks" % (batch, world_env))
        per_gpu = batch // max(world_env, 1)
    else:
        per_gpu = batch
    if args.impl == "reference":
        return run_reference(args, task, per_gpu)
    if args.warmup < 3:
        args.warmup = 3

    import numpy as np
    import torch
    import torch.distributed as dist
    import pybullet_multigoal_gym_b200 as pmg
    from pybullet_multigoal_gym_b200.sharded import ShardedKukaEnv

    rank = int(os.environ.get("RANK", "0"))
    world = world_env
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    distributed = world > 1
    if distributed:
        dist.init_process_group("nccl", device_id=dev)
    B, K, W = per_gpu, args.steps, args.warmup

    import contextlib
    import io
    kw = dict(num_block=4, check_actions=False, device_sampling=True, auto_reset=True, seed=1234)
    with contextlib.redirect_stdout(io.StringIO()):  # make_env prints 'Task id: ...' like the reference
        if distributed:
            env_s = ShardedKukaEnv(task, B * world, device=local_rank, fused=(None if args.gather == "fused" else False), **kw)  # None: peer-memory gather when the box allows, else NCCL on all ranks
            env = env_s.env
        else:
            env_s = None
            env = pmg.make_env(task=task, batch=B, device=local_rank, **kw)
    A, Wd = env.action_dim, env.row_width

    # staggered episodes: environment i (global index) starts at elapsed step i mod 50
    st = env.get_state()
    st[:, -1] = (np.arange(B) + rank * B) % EPISODE
    env.set_state(st)

    # synthetic random policy, resident in HBM before the timed region
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    n_tape = SETTLE_STEPS + W + max(K, E2E_STEPS)
    tape = torch.rand((n_tape, B, A), device=dev, generator=gen) * 2 - 1
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

    for t in range(SETTLE_STEPS + W):  # set-up (stationary contact mix) + warm-up, untimed
        one_step(t)
    torch.cuda.synchronize()
    if distributed:
        dist.barrier()
    launches0 = env.launch_count
    env.kernel_timing(True)
    sampler = ClockSampler(local_rank)
    sampler.start()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    torch.cuda.synchronize()
    for k in range(K):
        flush.zero_()  # L2 flush, outside the timed events
        starts[k].record()
        one_step(SETTLE_STEPS + W + k)  # step kernel + auto-reset pass (+ gather over peer memory)
        stops[k].record()
    torch.cuda.synchronize()
    if distributed:
        dist.barrier()
    sampler.stop_flag = True
    sampler.join()
    per_step = [s.elapsed_time(e) for s, e in zip(starts, stops)]
    total_ms = sum(per_step)
    kernel_total_ms, kernel_n = env.kernel_time_ms()
    env.kernel_timing(False)
    kernel_ms = kernel_total_ms / max(kernel_n, 1)
    launches = env.launch_count - launches0
    overflow = env.overflow_count
    t_ms = torch.tensor([total_ms, kernel_ms], device=dev, dtype=torch.float64)
    t_all = [torch.zeros_like(t_ms) for _ in range(world)] if distributed else [t_ms]
    if distributed:
        dist.all_gather(t_all, t_ms)
    total_ms_max = max(float(t[0]) for t in t_all)
    value = B * world * K / (total_ms_max / 1e3)

    # ---- e2e: public API, host buffers, H2D + kernel(s) + D2H every step --------------------------------------
    n_e2e = E2E_STEPS  # its own fixed length (reported as e2e.steps): the slowest environment's tail differs from step to step,
    # and 20 steps of it read 8 % off the 50-step figure
    host_tape = tape[SETTLE_STEPS + W:SETTLE_STEPS + W + n_e2e].cpu().numpy()
    stepper = (lambda a: env_s.step_host(a)) if env_s is not None else (lambda a: env.step(a))
    warm_tape = tape[SETTLE_STEPS:SETTLE_STEPS + W].cpu().numpy()
    for k in range(W):  # W warm-up steps as in the device-timed leg (pinned staging buffers, first-call paths), on random
        stepper(warm_tape[k])  # actions of their own: repeating one action would press more arms onto the table
    torch.cuda.synchronize()
    if distributed:
        dist.barrier()
    t0 = time.perf_counter()
    for k in range(n_e2e):
        stepper(host_tape[k])
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e_t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if distributed:
        dist.all_reduce(e_t, op=dist.ReduceOp.MAX)
    e2e_value = B * world * n_e2e / float(e_t.item())

    if rank == 0:
        hbm_peak, peak_src = peaks()
        clocks = sampler.summary()
        sm_hz = 1e6 * float(clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965.0)
        achieved = BYTES_PER_ENV_STEP[task] * B / (kernel_ms / 1e3) / 1e9
        ncu = ncu_summary(task, B)
        roof = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": ncu.get("dram_bytes_per_launch"), "peak_source": peak_src,
                "bytes_per_env_step": BYTES_PER_ENV_STEP[task], "kernel_ms": kernel_ms, "kernel_launches_timed": int(kernel_n),
                "note": ("not the binding roofline: 100 dependent substeps per env-step on ~0.3-1.1 KB of compulsory HBM traffic "
                         "(SURVEY.md 0.5); the kernel is bound by warp-instruction issue and dependent-issue latency")}
        if ncu.get("warp_inst_per_launch"):
            roof["issue_slot_frac"] = ncu["warp_inst_per_launch"] / (148 * 4 * sm_hz * kernel_ms / 1e3)
            roof["issue_slot_note"] = "ncu-measured warp instructions per launch / (148 SMs x 4 schedulers x measured SM clock x live kernel time)"
        if ncu.get("fp32_flop_per_launch"):
            roof["fp32_frac"] = ncu["fp32_flop_per_launch"] / (kernel_ms / 1e3) / (148 * 128 * 2 * sm_hz)
            roof["fp32_note"] = "ncu-measured fadd + fmul + 2 x ffma thread instructions per launch vs 148 x 128 lanes x 2 x SM clock"
        if ncu:
            roof["ncu_source"] = ncu.get("source")
        if env_s is not None and env_s.fused:
            par = "env-sharded x%d, gather fused into the step over peer memory (NVLink P2P stores + per-rank sequence flags), no NCCL on the data path" % world
        elif env_s is not None:
            par = "env-sharded x%d, one NCCL all-gather of the obs batch per step" % world
        else:
            par = "single GPU"
        kname = {"reach": "PMG_COOP", "push": "PMG_COOP_BLOCK", "pick_and_place": "PMG_COOP_BLOCK", "slide": "PMG_COOP_BLOCK"}.get(task, "PMG_COOP_STACK")
        coop = os.environ.get(kname, "1") != "0"
        d2h = (B * world if env_s is not None else B) * (Wd * 4 + 4 + 2)
        line = {
            "metric": metric_name(task, B), "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": total_ms_max / K, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_text(task, B, world, args.scaling), "global_batch": B * world, "parallelism": par,
                       "l2": "flushed between timed steps (256 MiB memset outside the per-step CUDA events)",
                       "kernel": ("lane-cooperative step kernel: 8 lanes per env, 4 envs per one-warp block" if coop
                                  else "thread-per-env step kernel: 32 envs per warp"),
                       "resets_per_step": B // EPISODE, "contact_pool_overflows": overflow},
            "roofline": roof,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * A * 4, "d2h_bytes_per_step": d2h, "steps": n_e2e,
                    "api": ("ShardedKukaEnv.step_host(numpy): H2D local actions, step + gather, D2H global batch" if env_s is not None
                            else "env.step(numpy) -> pmg_step_host_blocks")},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "breakdown": {"step_kernel_ms_per_rank": [float(t[1]) for t in t_all],
                          "step_total_ms_per_rank": [float(t[0]) / K for t in t_all],
                          "note": "total - kernel = auto-reset pass" + (" + peer push / wait for the slowest rank" if world > 1 else "")},
        }
        if not args.no_cpu_baseline and world == 1:
            threads = os.cpu_count() or 1
            n_env = cpu_sample_size(task, threads)
            n_steps = 100 if task in ("reach", "push", "pick_and_place", "slide") else 50
            rate, secs, flags = cpu_port_rate(task, n_env, n_steps, threads, 60)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": "%d of %d envs x %d steps (%.1f s) after 60 warm-up steps, staggered episodes with resets, %d pthreads, "
                                              "double-precision C port of the path built %s; the reference's pybullet backend is not installable "
                                              "here or on the GPU box" % (n_env, B, n_steps, secs, threads, flags)}
        print(json.dumps(line), flush=True)
    if distributed:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
