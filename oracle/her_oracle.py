"""numpy restatement of pmg_her_sample / pmg_her_relabel -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

The reference repository contains no relabelling code (its agents live in the author's drl_implementation
repo, README.md:18-20), so there is nothing on disk to pin this against: the oracle restates the published
"future" strategy of Hindsight Experience Replay (Andrychowicz et al. 2017, section 4.5: replay with goals
achieved later in the same episode) around the reference's own reward function
(kuka_single_step_base_env.py:237-244), and the counter-based sampler of the kernel bit-exactly."""
import numpy as np

_G = np.uint64(0x9e3779b97f4a7c15)


def _mix(z):
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xbf58476d1ce4e5b9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94d049bb133111eb)
    return z ^ (z >> np.uint64(31))


def sample(n, n_episodes, horizon, her_prob, seed):
    with np.errstate(over="ignore"):
        i = np.arange(n, dtype=np.uint64)
        base = np.uint64(seed) + _G * (np.uint64(3) * i + np.uint64(1))
        h0, h1, h2 = _mix(base), _mix(base + _G), _mix(base + np.uint64(2) * _G)
        ep = ((h0 >> np.uint64(32)) * np.uint64(n_episodes)) >> np.uint64(32)
        t = ((h1 >> np.uint64(32)) * np.uint64(horizon)) >> np.uint64(32)
        u = (h2 >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / 16777216.0)
        span = np.uint64(horizon) - t
        fut = t + np.uint64(1) + (((h2 & np.uint64(0xffffffff)) * span) >> np.uint64(32))
    fut = np.where(u < np.float32(her_prob), fut.astype(np.int64), -1)
    return ep.astype(np.int32), t.astype(np.int32), fut.astype(np.int32)


def relabel(ag, dg, ep, t, fut, thr=0.05, binary=True):
    """ag [E, T + 1, G], dg [E, G] -> goals [n, G], reward [n] (float32 sparse / float64 dense), achieved [n]."""
    ag = np.asarray(ag, dtype=np.float64)
    dg = np.asarray(dg, dtype=np.float64)
    goals = np.where((fut >= 0)[:, None], ag[ep, np.maximum(fut, 0)], dg[ep])
    d = np.linalg.norm(ag[ep, t + 1] - goals, axis=-1)     # _compute_reward
    not_achieved = d > thr
    reward = -not_achieved.astype(np.float32) if binary else -d
    return goals, reward, ~not_achieved
