"""Short rollouts of every task at an odd batch (a partial state tile, idle octets) with the jaws driven down onto the
table / the blocks, for compute-sanitizer (memcheck / racecheck) runs over the lane-cooperative kernels, the reset
kernels (host- and device-sampled, auto-reset) and the thread-per-env kernels (PMG_COOP*=0)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pybullet_multigoal_gym_b200 as pmg

B = int(sys.argv[1]) if len(sys.argv) > 1 else 37
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
for task in ("reach", "push", "pick_and_place", "slide", "block_stack", "block_rearrange"):
    if task == "slide" and os.environ.get("PMG_COOP_BLOCK") == "0":
        continue
    for dev_reset in (False, True):
        env = pmg.make_env(task=task, batch=B, num_block=3, check_actions=False, device_sampling=dev_reset, auto_reset=dev_reset, max_episode_steps=4)
        env.reset()
        gen = torch.Generator(device="cuda"); gen.manual_seed(3)
        for t in range(steps):
            a = torch.rand((B, env.action_dim), device="cuda", generator=gen) * 2 - 1
            a[:, 2] = -1.0
            obs, r, done, info = env.step(a)
            if not dev_reset and bool(done.any()):
                env.reset()
        torch.cuda.synchronize()
        print(task, "device reset" if dev_reset else "host reset", "finite:", bool(torch.isfinite(obs["observation"]).all()), "overflow", env.overflow_count, flush=True)
        env.close()
