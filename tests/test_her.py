"""Hindsight relabelling around _compute_reward: numpy oracle properties on the CPU,
bit-exact kernel-vs-oracle parity on the GPU (integer indices and gathered goals bit-exact, rewards identical
away from the distance threshold)."""
import numpy as np
import pytest

from oracle import her_oracle as H


def _episodes(E, T, G, seed):
    rng = np.random.RandomState(seed)
    ag = np.cumsum(rng.uniform(-0.01, 0.01, size=(E, T + 1, G)), axis=1).astype(np.float32) + rng.uniform(-0.1, 0.1, size=(E, 1, G)).astype(np.float32)
    dg = rng.uniform(-0.15, 0.15, size=(E, G)).astype(np.float32)
    return ag, dg


def test_oracle_sampler_is_the_future_strategy():
    E, T = 37, 50
    ep, t, fut = H.sample(100000, E, T, 0.8, seed=3)
    assert ep.min() == 0 and ep.max() == E - 1 and t.min() == 0 and t.max() == T - 1
    rel = fut >= 0
    assert abs(rel.mean() - 0.8) < 0.01                       # k = 4 future goals per real one
    assert np.all(fut[rel] > t[rel]) and fut.max() == T        # strictly later in the same episode, last index reachable
    # uniform over the remaining steps: the mean offset of t = 0 samples is (T + 1) / 2
    first = rel & (t == 0)
    assert abs(fut[first].mean() - (T + 1) / 2) < 1.5
    assert not np.array_equal(H.sample(1000, E, T, 0.8, seed=4)[0], ep[:1000])


def test_oracle_relabel_matches_reference_reward_semantics():
    ag, dg = _episodes(8, 20, 3, 0)
    ep, t, fut = H.sample(500, 8, 20, 0.8, seed=1)
    goals, r, ok = H.relabel(ag, dg, ep, t, fut, thr=0.05, binary=True)
    assert r.dtype == np.float32 and set(np.unique(r)) <= {-1.0, 0.0}   # sparse: -(d > thr) as float32
    keep = fut < 0
    assert np.array_equal(goals[keep], dg[ep[keep]].astype(np.float64))
    # a transition relabelled with its own next achieved goal is always a success
    same = fut == t + 1
    assert same.any() and ok[same].all() and np.all(r[same] == 0.0)
    _, rd, _ = H.relabel(ag, dg, ep, t, fut, thr=0.05, binary=False)
    assert rd.dtype == np.float64 and np.all(rd <= 0) and np.array_equal(rd < -0.05, ~ok)


@pytest.mark.gpu
@pytest.mark.parametrize("G,binary", [(3, True), (12, False)])
def test_her_kernels_match_oracle(G, binary):
    import torch
    from pybullet_multigoal_gym_b200 import her
    E, T, n = 64, 50, 20000
    ag, dg = _episodes(E, T, G, 5)
    ep, t, fut = her.sample(n, E, T, her_prob=0.8, seed=11)
    oe, ot, of = H.sample(n, E, T, 0.8, 11)
    assert np.array_equal(ep.cpu().numpy(), oe) and np.array_equal(t.cpu().numpy(), ot) and np.array_equal(fut.cpu().numpy(), of)
    goals, r, ok = her.relabel(torch.from_numpy(ag).cuda(), torch.from_numpy(dg).cuda(), ep, t, fut, 0.05, binary)
    og, orr, ook = H.relabel(ag, dg, oe, ot, of, 0.05, binary)
    assert np.array_equal(goals.cpu().numpy(), og.astype(np.float32))      # a gather: bit-exact
    d = np.linalg.norm(ag[oe, ot + 1].astype(np.float64) - og, axis=-1)
    clear = np.abs(d - 0.05) > 1e-6                                         # fp32 vs fp64 distance at the threshold
    assert np.array_equal(ok.cpu().numpy()[clear], ook[clear])
    if binary:
        assert np.array_equal(r.cpu().numpy()[clear], orr[clear])
    else:
        np.testing.assert_allclose(r.cpu().numpy(), orr, atol=1e-6)
