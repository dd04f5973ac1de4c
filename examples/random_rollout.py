#!/usr/bin/env python3
"""Random-policy rollout with the reference's API, batched on one B200 (cf. the reference's
examples/kuka_reach.py loop: env.reset(); env.step(env.action_space.sample()))."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import pybullet_multigoal_gym_b200 as pmg  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--task", default="reach", choices=["reach", "push", "pick_and_place", "block_stack"])
ap.add_argument("--batch", type=int, default=8192)
ap.add_argument("--episodes", type=int, default=2)
args = ap.parse_args()

env = pmg.make_env(task=args.task, gripper="parallel_jaw", num_block=4, render=False, binary_reward=True,
                   max_episode_steps=50, batch=args.batch)
for ep in range(args.episodes):
    obs = env.reset()
    t0 = time.perf_counter()
    done = torch.zeros(args.batch, dtype=torch.bool, device="cuda")
    steps = 0
    while not bool(done.all()):
        action = torch.rand((args.batch, env.action_dim), device="cuda") * 2 - 1
        obs, reward, done, info = env.step(action)
        steps += 1
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("episode %d: %d steps x %d envs in %.3f s (%.2f M env-steps/s), success rate %.3f" % (
        ep, steps, args.batch, dt, steps * args.batch / dt / 1e6, float(info["goal_achieved"].float().mean())))
