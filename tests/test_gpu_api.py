"""GPU tests of the env API surface: partial resets, seeding, determinism, TimeLimit, sharded wrapper."""
import contextlib
import io

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mk(task, batch, **kw):
    import pybullet_multigoal_gym_b200 as pmg
    with contextlib.redirect_stdout(io.StringIO()):
        return pmg.make_env(task=task, batch=batch, num_block=kw.pop("num_block", 4), **kw)


def test_partial_reset_only_touches_masked_envs():
    B = 16
    env = _mk("push", B)
    env.reset()
    a = torch.rand((B, 3), device="cuda") * 2 - 1
    for _ in range(3):
        obs, r, d, info = env.step(a)
    before = env.get_state()
    mask = np.zeros(B, dtype=bool)
    mask[[1, 5, 6]] = True
    obs2 = env.reset(mask=mask)
    after = env.get_state()
    keep = ~mask
    assert np.array_equal(before[keep], after[keep])
    assert np.all(after[mask, -1] == 0) and np.all(after[keep, -1] == 3)          # elapsed steps
    assert np.all(after[mask, 9:18] == 0)                                       # joint velocities zeroed
    assert not np.array_equal(before[mask, 46:49], after[mask, 46:49])          # block respawned
    for k in obs:  # rewritten by the reset kernel's own observation code: same numbers up to fp32 rounding
        assert torch.allclose(obs2[k][keep], obs[k][keep], rtol=0, atol=2e-6)
    # done follows the per-env counter
    for _ in range(47):
        obs, r, d, info = env.step(a)
    assert bool(d[keep].all()) and not bool(d[mask].any())


def test_seed_reproducibility_and_stream_offsets():
    e1, e2 = _mk("pick_and_place", 8, seed=11), _mk("pick_and_place", 8, seed=11)
    o1, o2 = e1.reset(), e2.reset()
    for k in o1:
        assert torch.equal(o1[k], o2[k])
    a = torch.rand((8, 4), device="cuda") * 2 - 1
    for _ in range(3):
        s1, s2 = e1.step(a), e2.step(a)
        for k in s1[0]:
            assert torch.equal(s1[0][k], s2[0][k])          # bitwise deterministic
    # env i of a batch seeded s follows the stream of seed s + i
    e3 = _mk("pick_and_place", 4, seed=13)
    e4 = _mk("pick_and_place", 8, seed=11)
    g3, g4 = e3.reset()["desired_goal"], e4.reset()["desired_goal"]
    assert torch.equal(g3[0], g4[2]) and torch.equal(g3[1], g4[3])
    assert e1.seed(5) == [5]
    with pytest.raises(ValueError):
        e1.seed(-3)


def test_time_limit_and_info_keys():
    env = _mk("reach", 4, max_episode_steps=3)
    env.reset()
    a = torch.zeros((4, 3), device="cuda")
    flags = []
    for _ in range(4):
        obs, r, d, info = env.step(a)
        flags.append(bool(d.all()))
        assert set(info) == {"goal_achieved", "is_success", "TimeLimit.truncated"}
        assert r.dtype == torch.float32 and set(np.unique(r.cpu().numpy())) <= {-1.0, 0.0}
    assert flags == [False, False, True, True]
    assert env._max_episode_steps == 3 and env.action_space.shape == (3,)
    assert env.observation_space["state"].shape == (4, 3)


def test_four_dim_action_on_reach_ignores_grip_column():
    e1, e2 = _mk("reach", 4), _mk("reach", 4)
    e1.reset()
    e2.reset()
    a4 = torch.rand((4, 4), device="cuda") * 2 - 1
    o1 = e1.step(a4)[0]
    o2 = e2.step(a4[:, :3].contiguous())[0]
    assert torch.equal(o1["observation"], o2["observation"])


def test_her_style_relabelling_loop():
    """The use the reference is built for: store an episode, relabel goals with achieved goals of
    later steps, recompute rewards with env._compute_reward on [N, G] batches."""
    B, T = 32, 10
    env = _mk("push", B, binary_reward=True)
    obs = env.reset()
    ag = [obs["achieved_goal"].clone()]
    for t in range(T):
        obs, r, d, info = env.step(torch.rand((B, 3), device="cuda") * 2 - 1)
        ag.append(obs["achieved_goal"].clone())
    ag = torch.stack(ag)                                   # [T+1, B, 3]
    t_idx = torch.randint(0, T, (64,), device="cuda")
    e_idx = torch.randint(0, B, (64,), device="cuda")
    f_idx = torch.minimum(t_idx + 1 + torch.randint(0, T, (64,), device="cuda"), torch.tensor(T, device="cuda"))
    new_goal = ag[f_idx, e_idx]
    reward, ok = env._compute_reward(ag[t_idx + 1, e_idx], new_goal)
    d = (ag[t_idx + 1, e_idx] - new_goal).norm(dim=-1)
    assert torch.equal(ok, ~(d > 0.05)) or bool(((d - 0.05).abs() < 1e-6).any())
    assert bool((reward[ok] == 0).all()) and bool((reward[~ok] == -1).all())
    same = env._compute_reward(new_goal, new_goal)
    assert bool(same[1].all()) and bool((same[0] == 0).all())


@pytest.mark.parametrize("task", ["reach", "push"])
def test_ragged_batches_are_independent_of_the_launch_geometry(task):
    """An environment's trajectory depends on its seed and actions only, not on how many neighbours share its
    warp / octet: batches of 1, 3, 5 and 33 environments (partial octets, partial warps, partial blocks of the
    cooperative and of the thread-per-env kernels) reproduce the first rows of a 64-environment batch bit-exactly."""
    import contextlib
    import io
    import pybullet_multigoal_gym_b200 as pmg

    def mk(b):
        with contextlib.redirect_stdout(io.StringIO()):
            return pmg.make_env(task=task, batch=b, check_actions=False)

    big = mk(64)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(9)
    acts = torch.rand((12, 64, big.action_dim), device="cuda", generator=gen) * 2 - 1
    acts[:8, :, 2] = -1.0   # down onto the table / block: contact paths included
    ref_reset = big.reset()
    ref = [big.step(acts[t]) for t in range(12)]
    for b in (1, 3, 5, 33):
        env = mk(b)
        o = env.reset()
        for k in o:
            assert torch.equal(o[k], ref_reset[k][:b]), (b, k)
        for t in range(12):
            obs, r, done, info = env.step(acts[t, :b].contiguous())
            for k in obs:
                assert torch.equal(obs[k], ref[t][0][k][:b]), (b, t, k)
            assert torch.equal(r, ref[t][1][:b]) and torch.equal(done, ref[t][2][:b])
            assert torch.equal(info["goal_achieved"], ref[t][3]["goal_achieved"][:b])
        assert env.overflow_count == 0


def test_both_host_buffer_entry_points_agree():
    """pmg_step_host (packed rows, the entry INTEGRATION.md's stub binds) and pmg_step_host_blocks (what env.step
    uses) deliver the same numbers."""
    import ctypes as C
    from pybullet_multigoal_gym_b200 import _lib
    B = 33
    e1, e2 = _mk("push", B), _mk("push", B)
    e1.reset()
    e2.reset()
    L = _lib.load()
    rng = np.random.RandomState(1)
    packed = np.zeros((B, e1.row_width), np.float32)
    r = np.zeros(B, np.float32)
    d, s = np.zeros(B, np.uint8), np.zeros(B, np.uint8)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    for t in range(3):
        a = rng.uniform(-1, 1, size=(B, 3)).astype(np.float32)
        _lib.check(L.pmg_step_host(e1._h, vp(a), vp(packed), vp(r), vp(d), vp(s), None))
        obs, r2, d2, info = e2.step(a)
        O, P, G = e1.obs_dim, e1.policy_dim, e1.goal_dim
        assert np.array_equal(obs["observation"], packed[:, :O]) and np.array_equal(obs["policy_state"], packed[:, O:O + P])
        assert np.array_equal(obs["achieved_goal"], packed[:, O + P:O + P + G]) and np.array_equal(obs["desired_goal"], packed[:, O + P + G:])
        assert np.array_equal(r2.astype(np.float32), r) and np.array_equal(d2, d.astype(bool)) and np.array_equal(info["goal_achieved"], s.astype(bool))


def test_device_path_action_check_runs_in_the_kernel():
    """kuka.py:168 asserts action_space.contains(a).  On the device path the step kernel tests the action rows itself
    and raises a flag in mapped host memory (pmg_action_error): no host synchronisation per step; the error surfaces
    at the next step that finds the kernel finished, or on demand through env.action_error()."""
    import torch
    import pybullet_multigoal_gym_b200 as pmg
    for task, bad in (("reach", 1.5), ("pick_and_place", float("nan")), ("block_stack", -1.01)):
        env = pmg.make_env(task=task, batch=37, num_block=2)
        env.reset()
        a = torch.zeros((37, env.action_dim), device="cuda")
        env.step(a)
        assert not env.action_error()
        a[23, env.action_dim - 1] = bad
        env.step(a)                      # runs, flagged by the kernel
        torch.cuda.synchronize()
        with pytest.raises(pmg.ActionError):
            env.step(torch.zeros_like(a))
        assert not env.action_error()    # the query cleared it
        env.step(torch.zeros_like(a))
        with pytest.raises(ValueError):  # host-buffer path: checked before anything is launched
            env.step(np.full((37, env.action_dim), 2.0, np.float32))
        env.close()


@pytest.mark.parametrize("task", ["reach", "pick_and_place", "block_stack"])
def test_state_tile_layout_does_not_change_results(task, monkeypatch):
    """The persistent state / manifold arrays are tiled ([tile][word][env in tile], 4 environments per tile for the
    lane-cooperative kernels); PMG_STATE_TILE=0 restores the plain [word][env] arrays of round 1.  Same seeds, same
    actions, odd batch (a partial last tile): bit-identical rows, rewards and flags, and get_state round-trips."""
    import pybullet_multigoal_gym_b200 as pmg
    outs = []
    for plain in (False, True):
        if plain:
            monkeypatch.setenv("PMG_STATE_TILE", "0")
        env = pmg.make_env(task=task, batch=37, num_block=3, seed=5)
        env.reset()
        gen = torch.Generator(device="cuda")
        gen.manual_seed(11)
        rows = []
        for t in range(12):
            a = torch.rand((37, env.action_dim), device="cuda", generator=gen) * 2 - 1
            a[:, 2] = -1.0   # down onto the table / the blocks: contacts, manifolds in use
            obs, r, done, info = env.step(a)
            rows.append(torch.cat([obs[k] for k in ("observation", "policy_state", "achieved_goal", "desired_goal")] + [r[:, None]], dim=1).cpu().numpy())
        st = env.get_state()
        env.set_state(st)
        assert np.array_equal(env.get_state(), st)
        outs.append((np.stack(rows), st))
        env.close()
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
