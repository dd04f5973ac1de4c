"""Short multi-block rollout for compute-sanitizer (racecheck / memcheck) runs of the cooperative multi-block kernel."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pybullet_multigoal_gym_b200 as pmg
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 14
for task in ("block_stack", "block_rearrange"):
    env = pmg.make_env(task=task, batch=B, num_block=4, check_actions=False)
    env.reset()
    gen = torch.Generator(device="cuda"); gen.manual_seed(3)
    for t in range(steps):
        a = torch.rand((B, env.action_dim), device="cuda", generator=gen) * 2 - 1
        a[:, 2] = -1.0   # drive the jaws down onto the table / the blocks
        obs, r, done, info = env.step(a)
    torch.cuda.synchronize()
    print(task, "finite:", bool(torch.isfinite(obs["observation"]).all()), "overflow", env.overflow_count, flush=True)
    env.close()
