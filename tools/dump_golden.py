#!/usr/bin/env python3
"""Golden-trace hook for the first environment where the REAL pybullet + gym are importable
(SURVEY.md 8c).  Runs the unmodified reference for each in-scope task with seed 0 and the committed
action tapes of tests/golden/ref_plumbing_*.npz, and reports how far the CPU oracle is from it.
Here (no pybullet) it only says so; nothing in the tests depends on it."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    try:
        import pybullet  # noqa: F401
        import gym  # noqa: F401
        if "pybullet_shim" in getattr(pybullet, "__file__", ""):
            raise ImportError("that is the shim")
    except ImportError:
        print("real pybullet/gym not importable here: physics parity stays UNPINNED (DESIGN.md section 3)")
        return 0
    ref_root = os.environ.get("PMG_REFERENCE", "/root/reference")
    sys.path.insert(0, ref_root)
    import pybullet_multigoal_gym as ref
    keys = ("observation", "policy_state", "achieved_goal", "desired_goal")
    for name, kw in [("reach", dict(task="reach")), ("push", dict(task="push", binary_reward=False)),
                     ("pick_and_place", dict(task="pick_and_place")), ("block_stack", dict(task="block_stack", num_block=4))]:
        g = np.load(os.path.join(ROOT, "tests", "golden", "ref_plumbing_%s.npz" % name))
        env = ref.make_env(gripper="parallel_jaw", render=False, **kw)
        L, k, worst = int(g["episode_len"]), 0, 0.0
        trace = []
        for ep in range(g["reset_obs"].shape[0]):
            o = env.reset()
            trace.append(np.concatenate([np.ravel(o[key]) for key in keys]))
            for t in range(L):
                o, r, d, info = env.step(g["actions"][k])
                flat = np.concatenate([np.ravel(o[key]) for key in keys])
                worst = max(worst, float(np.abs(flat - g["step_obs"][k]).max()))
                trace.append(flat)
                k += 1
        np.save(os.path.join(ROOT, "tests", "golden", "real_pybullet_%s.npy" % name), np.array(trace))
        print("%-16s max |real pybullet - oracle-backed golden| over %d steps: %.3g" % (name, k, worst))
    return 0


if __name__ == "__main__":
    sys.exit(main())
