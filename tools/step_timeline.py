"""Per-step kernel time across an episode (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pybullet_multigoal_gym_b200 as pmg
task = sys.argv[1] if len(sys.argv) > 1 else "reach"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
env = pmg.make_env(task=task, batch=B, num_block=4, check_actions=False)
n = 100
acts = torch.rand((n, B, env.action_dim), device="cuda") * 2 - 1
out = torch.empty((B, env.row_width), device="cuda"); r = torch.empty((B,), device="cuda")
d = torch.empty((B,), dtype=torch.uint8, device="cuda"); s = torch.empty((B,), dtype=torch.uint8, device="cuda")
env.reset()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
ev[0].record()
for t in range(n):
    if t == 50:
        env.reset()
    env.step_packed(acts[t], out, r, d, s)
    ev[t + 1].record()
torch.cuda.synchronize()
ms = [ev[t].elapsed_time(ev[t + 1]) for t in range(n)]
print(task, "per-step ms:", " ".join("%.2f" % m for m in ms))
tipz = out[:, 2]
print("tip z min/mean", float(tipz.min()), float(tipz.mean()))
