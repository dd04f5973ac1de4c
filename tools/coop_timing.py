"""Development aid: per-phase cycle counts of the cooperative kernels for octets with / without contacts.
Usage: coop_timing.py [down] [task[:batch]]   (default reach:8192)

Build (here):  nvcc ... -DPMG_COOP_TIMING -o gpurun_out/libpmg_timing.so   (see tools/gpu_calls/gpu_timing.sh)
Run (GPU box): PMG_LIBRARY=pybullet_multigoal_gym_b200/libpmg_timing.so python tools/coop_timing.py"""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pybullet_multigoal_gym_b200 as pmg
from pybullet_multigoal_gym_b200 import _lib

args = sys.argv[1:]
down = "down" in args
args = [a for a in args if a != "down"]
task, _, bs = (args[0] if args else "reach:8192").partition(":")
B = int(bs or 8192)
env = pmg.make_env(task=task, batch=B, num_block=4, check_actions=False)
L = _lib.load()
torch.manual_seed(0)
acts = torch.rand((50, B, env.action_dim), device="cuda") * 2 - 1
if down:
    acts[:, :, 2] = -1.0  # every arm onto the table: all octets run the contact path together
out = torch.empty((B, env.row_width), device="cuda"); r = torch.empty((B,), device="cuda")
d = torch.empty((B,), dtype=torch.uint8, device="cuda"); s = torch.empty((B,), dtype=torch.uint8, device="cuda")
buf = (C.c_ulonglong * 16)()
L.pmg_debug_coop_cycles(buf)
for t in range(50):
    env.step_packed(acts[t], out, r, d, s)
    if t in (5, 30, 45):
        L.pmg_debug_coop_cycles(buf)
        c = list(buf)
        hot_n, cold_n = max(c[4], 1), max(c[6], 1)
        print("after step %2d: hot octet-substeps %8d: substep %7.0f cyc = narrowphase %6.0f + row set-up %6.0f + sweeps %6.0f + rest %6.0f | cold %9d: substep %6.0f cyc (broadphase %4.0f)"
              % (t, c[4], c[0] / hot_n, c[1] / hot_n, c[2] / hot_n, c[3] / hot_n, (c[0] - c[1] - c[2] - c[3]) / hot_n, c[6], c[5] / cold_n, c[7] / cold_n))
        if c[15]:
            print("               contact_sweep (one-block / Reach kernels): %.2f calls per hot substep, %.1f row visits per call; per call: normal loop %5.0f + friction loop %5.0f cycles of %5.0f (call, barriers, delta-velocity round trip: the rest)"
                  % (c[15] / hot_n, c[14] / c[15], c[12] / c[15], c[13] / c[15], c[3] / c[15]))
        if c[11]:
            print("               narrowphase of a touching pair: box_box %5.0f + manifold_add %5.0f + refresh %5.0f cycles" % (c[8] / c[11], c[9] / c[11], c[10] / c[11]))
