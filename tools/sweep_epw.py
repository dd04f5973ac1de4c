"""Sweep of PMG_ENVS_PER_WARP (development aid)."""
import sys, os, subprocess
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, os; sys.path.insert(0, %r)
import torch
import pybullet_multigoal_gym_b200 as pmg
task, B = sys.argv[1], int(sys.argv[2])
env = pmg.make_env(task=task, batch=B, num_block=4, check_actions=False)
acts = torch.rand((60, B, env.action_dim), device="cuda") * 2 - 1
out = torch.empty((B, env.row_width), device="cuda"); r = torch.empty((B,), device="cuda")
d = torch.empty((B,), dtype=torch.uint8, device="cuda"); s = torch.empty((B,), dtype=torch.uint8, device="cuda")
for t in range(10): env.step_packed(acts[t], out, r, d, s)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for t in range(50): env.step_packed(acts[10 + t], out, r, d, s)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 50
print("%%-16s B=%%6d epw=%%2s  %%.3f ms/step  %%.3f M env-steps/s" %% (task, B, os.environ.get("PMG_ENVS_PER_WARP", "auto"), ms, B / ms / 1e3), flush=True)
''' % root
for task, B in [("reach", 8192), ("push", 4096)]:
    for epw in ["32", "16", "8", "4"]:
        env = dict(os.environ)
        if epw != "auto":
            env["PMG_ENVS_PER_WARP"] = epw
        else:
            env.pop("PMG_ENVS_PER_WARP", None)
        out = subprocess.run([sys.executable, "-c", code, task, str(B)], env=env, capture_output=True, text=True)
        print("\n".join(l for l in out.stdout.splitlines() if "Task id" not in l) or out.stderr[-500:], flush=True)
