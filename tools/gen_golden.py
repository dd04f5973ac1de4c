#!/usr/bin/env python3
"""Generates tests/golden/*.npz by executing the reference's UNMODIFIED Python
(/root/reference/pybullet_multigoal_gym) on top of oracle/pybullet_shim.

Runs only in the build container (it imports /root/reference).  What the vectors pin: everything the
reference itself implements -- env-id plumbing, seeding, object/goal sampling streams, the action
map, motor commands, call order, observation layout, clipping, reward, TimeLimit -- because that code
runs for real.  What they do not pin: Bullet's arithmetic (the shim answers physics calls with our
CPU oracle), see DESIGN.md.
"""
import contextlib
import io
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "pybullet_shim"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

import pybullet_multigoal_gym as ref  # noqa: E402  (the reference package)

OUT = os.path.join(ROOT, "tests", "golden")
KEYS = ("observation", "policy_state", "achieved_goal", "desired_goal")
CONFIGS = [
    # two full 50-step episodes each, so that gym's TimeLimit flips `done` at the last step
    ("reach", dict(task="reach", binary_reward=True), 3, 100),
    ("push", dict(task="push", binary_reward=False), 3, 100),
    ("pick_and_place", dict(task="pick_and_place", binary_reward=True), 4, 100),
    ("block_stack", dict(task="block_stack", binary_reward=True, num_block=4), 4, 100),
    # Slide: long table, puck, goals beyond reach
    ("slide", dict(task="slide", binary_reward=False), 3, 100),
    # "next" rows of SURVEY.md 8(f): same physics, more of the reference's plumbing
    ("block_rearrange", dict(task="block_rearrange", binary_reward=True, num_block=3), 3, 100),
    ("block_stack_grip", dict(task="block_stack", binary_reward=True, num_block=3, grip_informed_goal=True), 4, 100),
    ("reach_jc", dict(task="reach", binary_reward=True, joint_control=True), 7, 100),
    ("pick_and_place_jc", dict(task="pick_and_place", binary_reward=False, joint_control=True), 8, 100),
    # task decomposition: env.set_sub_goal(k) is called by the script below (SUB_GOAL_SCHEDULE)
    ("block_stack_td", dict(task="block_stack", binary_reward=True, num_block=3, task_decomposition=True), 4, 100),
    ("block_stack_td_grip", dict(task="block_stack", binary_reward=True, num_block=3, task_decomposition=True, grip_informed_goal=True), 4, 100),
    # curriculum: 24 short episodes with updates activated and 4 goals per level, so that curriculum_prob walks
    # through its whole schedule (kuka_multi_step_base_env.py:350-379); max_episode_steps = 2
    ("block_stack_cur", dict(task="block_stack", binary_reward=True, num_block=3, use_curriculum=True, num_goals_to_generate=12, max_episode_steps=2), 4, 48),
    ("block_stack_cur_grip", dict(task="block_stack", binary_reward=True, num_block=3, use_curriculum=True, grip_informed_goal=True, num_goals_to_generate=12, max_episode_steps=2), 4, 48),
    # block_rearrange with the curriculum (kuka_multi_step_envs.py:193-227): level + 1 randomly chosen blocks get targets
    ("block_rearrange_cur", dict(task="block_rearrange", binary_reward=True, num_block=3, use_curriculum=True, num_goals_to_generate=12, max_episode_steps=2), 3, 48),
]
# (step within the episode) -> sub-goal index handed to env.set_sub_goal before that step; indices beyond the
# variant's number of sub-goals are taken modulo it by the script
SUB_GOAL_SCHEDULE = {0: 0, 10: 1, 20: 2, 30: 5, 40: -1}
# OracleEnv / make_env keyword arguments that reproduce each golden (shared with the tests)
VARIANTS = {
    "reach": dict(task="reach"), "push": dict(task="push", binary_reward=False),
    "pick_and_place": dict(task="pick_and_place"), "block_stack": dict(task="block_stack", num_block=4),
    "slide": dict(task="slide", binary_reward=False),
    "block_rearrange": dict(task="block_rearrange", num_block=3),
    "block_stack_grip": dict(task="block_stack", num_block=3, grip_informed_goal=True),
    "reach_jc": dict(task="reach", joint_control=True),
    "pick_and_place_jc": dict(task="pick_and_place", binary_reward=False, joint_control=True),
    "block_stack_td": dict(task="block_stack", num_block=3, task_decomposition=True),
    "block_stack_td_grip": dict(task="block_stack", num_block=3, task_decomposition=True, grip_informed_goal=True),
    "block_stack_cur": dict(task="block_stack", num_block=3, use_curriculum=True, num_goals_to_generate=12, max_episode_steps=2),
    "block_stack_cur_grip": dict(task="block_stack", num_block=3, use_curriculum=True, grip_informed_goal=True, num_goals_to_generate=12, max_episode_steps=2),
    "block_rearrange_cur": dict(task="block_rearrange", num_block=3, use_curriculum=True, num_goals_to_generate=12, max_episode_steps=2),
}


def pack(obs):
    return np.concatenate([np.asarray(obs[k], dtype=np.float64).ravel() for k in KEYS])


def scripted_actions(name, adim, T, rng):
    """Random actions with a scripted bias towards the table / first block so that contacts occur."""
    a = rng.uniform(-1, 1, size=(T, adim))
    if name == "reach":
        a[:14, 2] = -1.0
    if name.endswith("_jc"):
        a[:, :7] *= 0.4      # joint-space deltas of up to 0.02 rad per step
        a[:12, 1] = 0.5      # lean the arm forward and down: jaws / block / table contacts
        a[:12, 3] = 0.3
    return a


def main():
    os.makedirs(OUT, exist_ok=True)
    only = set(sys.argv[1:])  # e.g. `gen_golden.py slide`: (re)generate just these
    for name, kw, adim, T in CONFIGS:
        if only and name not in only:
            continue
        # make_env registers an env id once per process and the id does not encode num_block /
        # grip_informed_goal (__init__.py:56-85: first registration wins), so start from a clean registry
        from gym.envs.registration import registry
        registry.env_specs.clear()
        with contextlib.redirect_stdout(io.StringIO()):
            env = ref.make_env(gripper="parallel_jaw", render=False, **kw)
        rng = np.random.RandomState(2024)
        actions = scripted_actions(name, adim, T, rng)
        resets, steps, rewards, dones, oks = [], [], [], [], []
        sub_goal_calls, sub_goal_returns = [], []
        cur_prob, cur_level, cur_goal_step, cur_moved = [], [], [], []
        L = T // 2
        if "_cur" in name:
            L = 2
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                env.activate_curriculum_update()
        for ep in range(T // L):
            resets.append(pack(env.reset()))
            if "_cur" in name:
                inner = env.unwrapped if hasattr(env, "unwrapped") else env.env
                cur_prob.append(np.array(inner.curriculum_prob, dtype=np.float64))
                if kw["task"] == "block_rearrange":  # the level is implied by the blocks it picked (:204-208)
                    cur_level.append(len(inner.last_ind_block_to_move) - 1)
                    cur_moved.append(sum(1 << int(b) for b in inner.last_ind_block_to_move))
                else:
                    cur_level.append(int(inner.last_curriculum_level))
                cur_goal_step.append(int(inner.curriculum_goal_step))
            for t in range(L):
                a = actions[ep * L + t]
                if name not in ("reach", "reach_jc", "pick_and_place_jc"):
                    # steer the tip to the (first) block: push from the side / descend with open jaws
                    obs_now = steps[-1] if (steps and t > 0) else resets[-1]
                    tip = obs_now[0:3]
                    blk = obs_now[3:6] if not name.startswith("block_") else obs_now[8:11]
                    tgt = blk + np.array([0.0, 0.0, 0.0 if (name in ("push", "slide", "block_rearrange") or t > 10) else 0.07])
                    a[:3] = np.clip((tgt - tip) / 0.01, -1, 1)
                    if adim == 4:
                        a[3] = -1.0 if t < 16 else 1.0
                    actions[ep * L + t] = a
                if "_td" in name and t in SUB_GOAL_SCHEDULE:
                    nsub = 6 if name.endswith("grip") else 3
                    ind = SUB_GOAL_SCHEDULE[t]
                    ind = ind if ind < 0 else ind % nsub
                    sub = env.set_sub_goal(ind)
                    sub_goal_calls.append((ep * L + t, ind))
                    sub_goal_returns.append(np.asarray(sub, dtype=np.float64))
                obs, r, done, info = env.step(a.astype(np.float64))
                if t == L - 1:
                    assert done and info.get("TimeLimit.truncated") is True
                steps.append(pack(obs))
                rewards.append(float(r))
                dones.append(bool(done))
                oks.append(bool(info["goal_achieved"]))
        np.savez_compressed(os.path.join(OUT, "ref_plumbing_%s.npz" % name), actions=actions, reset_obs=np.array(resets),
                            step_obs=np.array(steps), reward=np.array(rewards), done=np.array(dones),
                            goal_achieved=np.array(oks), episode_len=L,
                            curriculum_prob=np.array(cur_prob), curriculum_level=np.array(cur_level, dtype=np.int64),
                            curriculum_goal_step=np.array(cur_goal_step, dtype=np.int64),
                            curriculum_moved_mask=np.array(cur_moved, dtype=np.int64),
                            dims=np.array([len(np.ravel(obs[k])) for k in KEYS]),
                            max_episode_steps=env._max_episode_steps,
                            sub_goal_calls=np.array(sub_goal_calls, dtype=np.int64).reshape(-1, 2),
                            sub_goal_returns=np.array(sub_goal_returns))
        print("%-16s reset %s steps %s  successes %d  last reward %s" % (name, np.array(resets).shape, np.array(steps).shape, sum(oks), rewards[-1]))


if __name__ == "__main__":
    main()
