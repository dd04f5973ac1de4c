"""Measures the two pieces either side of the step kernel that VERDICT r01 found unmeasured:
  (1) hindsight relabelling (pmg_her_sample + pmg_her_relabel): achieved bandwidth against the measured HBM copy
      bandwidth -- algorithmic bytes per sample (3 indices read + goal gather G + achieved goal G + goal out G + reward)
      = (3 G + 4) * 4 + 1; the episode store is larger than L2 so the gathers come from HBM;
  (2) env.reset() latency at the configs' batches: host MT19937 sampling (reference stream: one host thread draws every
      environment's rows, uploads them, reset kernel) against device Philox sampling (reset kernel only), and the
      masked auto-reset pass behind a step (B/50 environments)."""
import json
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import pybullet_multigoal_gym_b200 as pmg
from pybullet_multigoal_gym_b200 import her

peak = 6542.4
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
dev = torch.device("cuda:0")
print("== hindsight relabelling")
for G, E, T, n in [(3, 1 << 18, 50, 1 << 24), (12, 1 << 16, 50, 1 << 24), (3, 8192, 50, 1 << 20)]:
    ag = torch.rand((E, T + 1, G), device=dev); dg = torch.rand((E, G), device=dev)
    ep, t, fut = her.sample(n, E, T, her_prob=0.8, seed=1)
    for _ in range(2):
        her.relabel(ag, dg, ep, t, fut)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        goals, r, ok = her.relabel(ag, dg, ep, t, fut)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    byts = n * ((3 * G + 4) * 4 + 1)
    e0.record()
    for _ in range(reps):
        her.sample(n, E, T, her_prob=0.8, seed=1)
    e1.record(); torch.cuda.synchronize()
    ms_s = e0.elapsed_time(e1) / reps
    print("G=%2d episodes=%7d (store %.2f GB) samples=%9d: relabel %.3f ms = %.0f GB/s algorithmic = %.1f %% of the measured %.0f GB/s "
          "(random 4G-byte gathers: a 32-byte sector is fetched per %d-byte goal); sample %.3f ms = %.0f GB/s of index writes"
          % (G, E, ag.numel() * 4 / 1e9, n, ms, byts / ms / 1e6, 100 * byts / ms / 1e6 / peak, peak, 4 * G, ms_s, n * 12 / ms_s / 1e6))
    del ag, dg, ep, t, fut, goals, r, ok

print("== reset latency (wall clock around env.reset(), synchronised)")
for task, B in [("reach", 8192), ("push", 4096), ("pick_and_place", 4096), ("block_stack", 2048)]:
    for mode in ("host MT19937 (reference stream)", "device Philox"):
        env = pmg.make_env(task=task, batch=B, num_block=4, check_actions=False, device_sampling=mode.startswith("device"))
        env.reset(device_output=True)
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            t0 = time.perf_counter()
            env.reset(device_output=True)
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        print("%-15s B=%5d  %-32s full reset %.3f ms (median of 5)" % (task, B, mode, 1e3 * float(np.median(ts))), flush=True)
        env.close()
