"""Hindsight relabelling on the device, around `_compute_reward`.

The reference's agents (README.md:18-20, the author's drl_implementation repo) store whole episodes and, when
sampling a minibatch, replace the desired goal of most transitions by a goal achieved later in the same episode
("future" strategy), then call `env._compute_reward(achieved_goal, new_goal)`
(kuka_single_step_base_env.py:237-244).  With thousands of environments stepping on the GPU the episodes never
need to leave it: `sample` draws (episode, t, future) triples, `relabel` gathers the goals and evaluates the
reward in one pass.  Both are thin wrappers over the C-ABI (pmg_her_sample / pmg_her_relabel)."""
import ctypes as C

import torch

from . import _lib


def _ptr(t):
    return C.c_void_p(t.data_ptr())


def sample(n, n_episodes, horizon, her_prob=0.8, seed=0, device=0):
    """-> (episode, t, future) int32 CUDA tensors of length n; future = -1 keeps the original goal."""
    L = _lib.load()
    dev = torch.device("cuda", device)
    with torch.cuda.device(dev):
        ep = torch.empty((n,), dtype=torch.int32, device=dev)
        t = torch.empty((n,), dtype=torch.int32, device=dev)
        fut = torch.empty((n,), dtype=torch.int32, device=dev)
        _lib.check(L.pmg_her_sample(n, n_episodes, horizon, her_prob, seed, _ptr(ep), _ptr(t), _ptr(fut),
                                    C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    return ep, t, fut


def relabel(achieved_goals, desired_goals, episode, t, future, distance_threshold=0.05, binary_reward=True):
    """achieved_goals [E, T + 1, G], desired_goals [E, G] (CUDA float32) -> (goals [n, G], reward [n], achieved [n])."""
    L = _lib.load()
    ag = achieved_goals.to(torch.float32).contiguous()
    dg = desired_goals.to(torch.float32).contiguous()
    E, T1, G = ag.shape
    if tuple(dg.shape) != (E, G):
        raise AssertionError("desired_goals must have shape (%d, %d)" % (E, G))
    n = int(episode.numel())
    dev = ag.device
    # the kernel reads int32 indices: convert (int64 tensors would be misread) and keep them on the goals' device
    episode, t, future = (x.to(device=dev, dtype=torch.int32).contiguous() for x in (episode, t, future))
    if not (t.numel() == n and future.numel() == n):
        raise ValueError("episode, t and future must have the same length")
    with torch.cuda.device(dev):
        goals = torch.empty((n, G), dtype=torch.float32, device=dev)
        reward = torch.empty((n,), dtype=torch.float32, device=dev)
        ok = torch.empty((n,), dtype=torch.uint8, device=dev)
        _lib.check(L.pmg_her_relabel(_ptr(ag), _ptr(dg), E, T1 - 1, G, _ptr(episode), _ptr(t), _ptr(future), n,
                                     distance_threshold, int(binary_reward), _ptr(goals), _ptr(reward), _ptr(ok),
                                     C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    return goals, reward, ok.bool()
