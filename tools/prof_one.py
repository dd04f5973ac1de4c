"""Runs a few env.step launches of one task (for ncu captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pybullet_multigoal_gym_b200 as pmg
task = sys.argv[1] if len(sys.argv) > 1 else "reach"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
n = int(sys.argv[3]) if len(sys.argv) > 3 else 4
env = pmg.make_env(task=task, batch=B, num_block=4, check_actions=False)
acts = torch.rand((n, B, env.action_dim), device="cuda") * 2 - 1
if len(sys.argv) > 4 and sys.argv[4] == "down":
    acts[:, :, 2] = -1.0  # drive every arm onto the table: all lanes take the contact path
out = torch.empty((B, env.row_width), device="cuda"); r = torch.empty((B,), device="cuda")
d = torch.empty((B,), dtype=torch.uint8, device="cuda"); s = torch.empty((B,), dtype=torch.uint8, device="cuda")
for t in range(n):
    env.step_packed(acts[t], out, r, d, s)
torch.cuda.synchronize()
