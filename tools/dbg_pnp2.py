import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pybullet_multigoal_gym_b200 as pmg
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
env = pmg.make_env(task="pick_and_place", batch=B, max_episode_steps=80)
obs = env.reset()
phase = torch.zeros(B, dtype=torch.long, device="cuda"); timer = torch.zeros(B, dtype=torch.long, device="cuda")
for t in range(80):
    tip, blk, goal = obs["observation"][:, 0:3], obs["achieved_goal"], obs["desired_goal"]
    hover = blk + torch.tensor([0.0, 0.0, 0.06], device="cuda")
    tgt = torch.where((phase == 0)[:, None], hover, blk)
    tgt = torch.where((phase == 3)[:, None], goal, tgt)
    a = torch.zeros((B, 4), device="cuda")
    a[:, :3] = torch.clamp((tgt - tip) / 0.01, -1, 1)
    a[:, 3] = torch.where(phase >= 2, torch.ones(B, device="cuda"), -torch.ones(B, device="cuda"))
    hold = phase == 2
    a[hold, :3] = 0.0
    err = (tgt - tip).norm(dim=1)
    timer = torch.where(hold, timer + 1, timer)
    phase = torch.where((phase == 0) & (err < 0.008), torch.ones_like(phase), phase)
    phase = torch.where((phase == 1) & (err < 0.004), torch.full_like(phase, 2), phase)
    phase = torch.where((phase == 2) & (timer >= 4), torch.full_like(phase, 3), phase)
    obs, r, done, info = env.step(a)
    if t % 10 == 9:
        print(t, "phases", torch.bincount(phase, minlength=4).tolist(), "succ", float(info["goal_achieved"].float().mean()), "blk z mean", float(obs["achieved_goal"][:,2].mean()))
ok = info["goal_achieved"]
print("success", float(ok.float().mean()), "first 32:", ok[:32].int().tolist())
bad = (~ok).nonzero().flatten()[:6].tolist()
for i in bad:
    print(i, "phase", int(phase[i]), "tip", obs["observation"][i,:3].tolist(), "blk", obs["achieved_goal"][i].tolist(), "goal", obs["desired_goal"][i].tolist(), "close", float(obs["observation"][i,6]))
