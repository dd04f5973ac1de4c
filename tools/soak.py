"""Random-policy soak: several episodes per task, checks for NaN/inf, runaway states and pool overflows."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pybullet_multigoal_gym_b200 as pmg
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
episodes = int(sys.argv[2]) if len(sys.argv) > 2 else 3
CASES = [("reach", {}), ("push", {}), ("pick_and_place", {}), ("block_stack", {}), ("block_rearrange", {}),
         ("block_stack", {"grip_informed_goal": True}), ("reach", {"joint_control": True}), ("pick_and_place", {"joint_control": True})]
for task, extra in CASES:
    env = pmg.make_env(task=task, batch=B, num_block=4, check_actions=False, **extra)
    gen = torch.Generator(device="cuda"); gen.manual_seed(7)
    bad = 0; succ = 0; maxabs = 0.0
    for ep in range(episodes):
        env.reset()
        for t in range(50):
            a = torch.rand((B, env.action_dim), device="cuda", generator=gen) * 2 - 1
            if extra.get("joint_control"):
                a[:, :7] *= 0.3
            obs, r, done, info = env.step(a)
            o = obs["observation"]
            bad += int((~torch.isfinite(o)).any(dim=1).sum())
            maxabs = max(maxabs, float(o.abs().max()))
        succ += int(info["goal_achieved"].sum())
        assert bool(done.all())
    zmin = float(obs["achieved_goal"][:, :3 * max(1, env.num_block)].reshape(B, -1, 3)[..., 2].min())
    print("%-15s %-28s B=%d episodes=%d non-finite rows=%d max|obs|=%.3f min achieved z=%.4f successes(last step)=%d overflow=%d" % (
        task, extra, B, episodes, bad, maxabs, zmin, succ, env.overflow_count), flush=True)
    env.close()
