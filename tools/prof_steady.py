"""Brings one task to the steady state bench.py measures (staggered episodes, device auto-reset, 60 set-up steps) and
runs a few more steps -- the launches ncu captures (skip the first 60 step-kernel launches)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import pybullet_multigoal_gym_b200 as pmg
task = sys.argv[1] if len(sys.argv) > 1 else "reach"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
n = int(sys.argv[3]) if len(sys.argv) > 3 else 4
env = pmg.make_env(task=task, batch=B, num_block=4, check_actions=False, device_sampling=True, auto_reset=True, seed=1234)
st = env.get_state()
st[:, -1] = np.arange(B) % 50
env.set_state(st)
gen = torch.Generator(device="cuda"); gen.manual_seed(1234)
acts = torch.rand((60 + n, B, env.action_dim), device="cuda", generator=gen) * 2 - 1
out = torch.empty((B, env.row_width), device="cuda"); r = torch.empty((B,), device="cuda")
d = torch.empty((B,), dtype=torch.uint8, device="cuda"); s = torch.empty((B,), dtype=torch.uint8, device="cuda")
for t in range(60 + n):
    env.step_packed(acts[t], out, r, d, s)
torch.cuda.synchronize()
