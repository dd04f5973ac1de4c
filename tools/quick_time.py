"""Quick device-side timing of env.step (development aid; bench.py is the contract)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pybullet_multigoal_gym_b200 as pmg

CASES = [("reach", 8192), ("push", 4096), ("pick_and_place", 4096), ("block_stack", 2048), ("reach", 65536)]
if len(sys.argv) > 1:  # e.g. quick_time.py reach:8192 reach:65536 pick_and_place:4096:jc
    CASES = [tuple(a.split(":")) for a in sys.argv[1:]]
for case in CASES:
    task, B, jc = case[0], int(case[1]), len(case) > 2 and case[2] == "jc"
    env = pmg.make_env(task=task, batch=B, num_block=4, check_actions=False, joint_control=jc)
    A = env.action_dim
    acts = torch.rand((60, B, A), device="cuda") * 2 - 1
    out = torch.empty((B, env.row_width), device="cuda"); r = torch.empty((B,), device="cuda")
    d = torch.empty((B,), dtype=torch.uint8, device="cuda"); s = torch.empty((B,), dtype=torch.uint8, device="cuda")
    for t in range(5):
        env.step_packed(acts[t], out, r, d, s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 45
    for t in range(n):
        env.step_packed(acts[5 + t], out, r, d, s)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print("%-16s" % (task + ("_jc" if jc else "")) + " B=%6d  %.3f ms/step  %.3f M env-steps/s  overflow=%d" % (B, ms, B / ms / 1e3, env.overflow_count), flush=True)
    env.close()
