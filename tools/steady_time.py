"""Development aid: steady-state device time per env.step of several task:batch cases (the workload bench.py times:
staggered episodes, device auto-reset, 60 set-up steps), kernel time from the library's own events.
Usage: steady_time.py reach:8192 block_stack:2048 ...   (environment variables such as PMG_COOP_WPB apply)"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import pybullet_multigoal_gym_b200 as pmg

cases = [a.split(":") for a in sys.argv[1:]] or [["reach", "8192"]]
tag = " ".join("%s=%s" % (k, v) for k, v in sorted(os.environ.items()) if k.startswith("PMG_COOP"))
for task, bs in cases:
    B = int(bs)
    env = pmg.make_env(task=task, batch=B, num_block=4, check_actions=False, device_sampling=True, auto_reset=True, seed=1234)
    st = env.get_state()
    st[:, -1] = np.arange(B) % 50
    env.set_state(st)
    gen = torch.Generator(device="cuda"); gen.manual_seed(1234)
    n = 50
    acts = torch.rand((60 + n, B, env.action_dim), device="cuda", generator=gen) * 2 - 1
    out = torch.empty((B, env.row_width), device="cuda"); r = torch.empty((B,), device="cuda")
    d = torch.empty((B,), dtype=torch.uint8, device="cuda"); s = torch.empty((B,), dtype=torch.uint8, device="cuda")
    for t in range(60):
        env.step_packed(acts[t], out, r, d, s)
    torch.cuda.synchronize()
    env.kernel_timing(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(n):
        env.step_packed(acts[60 + t], out, r, d, s)
    e1.record(); torch.cuda.synchronize()
    k_ms, k_n = env.kernel_time_ms()
    env.kernel_timing(False)
    ms = e0.elapsed_time(e1) / n
    print("%-16s B=%6d  %.3f ms/step (step kernel %.3f)  %.3f M env-steps/s  overflow=%d  [%s]"
          % (task, B, ms, k_ms / max(k_n, 1), B / ms / 1e3, env.overflow_count, tag), flush=True)
    env.close()
