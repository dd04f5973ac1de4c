"""GPU tests of the device-side reset path (pmg_set_device_rng / pmg_reset_device / pmg_set_auto_reset): spawn rows
bit-exact against the numpy restatement of the sampler (oracle/device_rng_oracle.py), reset observations against the
CPU oracle placed on the same rows, masked resets with a device mask, and auto-reset against step + explicit reset."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

KEYS = ("observation", "policy_state", "achieved_goal", "desired_goal")
CASES = [("reach", {}), ("push", dict(binary_reward=False)), ("pick_and_place", {}), ("slide", {}), ("block_stack", dict(num_block=4)),
         ("block_stack", dict(num_block=3, grip_informed_goal=True)), ("block_rearrange", dict(num_block=4))]


def _mk(task, batch, **kw):
    import contextlib
    import io
    import pybullet_multigoal_gym_b200 as pmg
    with contextlib.redirect_stdout(io.StringIO()):
        return pmg.make_env(task=task, batch=batch, **kw)


def _np(t):
    return t.detach().cpu().numpy()


@pytest.mark.parametrize("task,kw", CASES)
def test_device_sampled_reset_matches_restatement_and_oracle(oracle, task, kw):
    from oracle import device_rng_oracle as R
    B, seed, base = 16, 77, 1000
    env = _mk(task, B, device_sampling=True, seed=seed, env_index_base=base, **kw)   # the ctor's reset is episode 0
    nb = env.num_block
    for episode in (1, 2):
        obs = env.reset()
        rows = env.last_spawn()
        for i in range(B):
            want = R.sample_row(task, nb, int(bool(kw.get("grip_informed_goal"))), seed, base + i, episode)
            assert np.array_equal(rows[i].view(np.uint32), want.view(np.uint32)), (task, i, episode, rows[i], want)
            o = oracle.OracleEnv(task, seed=0, **kw)
            o.reset()
            for _ in range(episode):  # the IK rest pose is refined by every reset (kuka.py:159)
                ref = o.reset_with(rows[i].astype(np.float64))
            for k in KEYS:
                np.testing.assert_allclose(_np(obs[k][i]), ref[k], atol=2e-6, err_msg="%s env %d %s" % (task, i, k))


def test_masked_device_reset_with_a_device_mask():
    B = 32
    env = _mk("push", B, device_sampling=True, seed=5)
    env.reset()
    before = env.last_spawn().copy()
    st0 = env.get_state()
    mask = torch.zeros(B, dtype=torch.bool, device="cuda")
    mask[3] = mask[17] = mask[31] = True
    env.step(torch.zeros((B, 3), device="cuda"))
    env.reset(mask=mask)
    after, st1 = env.last_spawn(), env.get_state()
    m = _np(mask)
    assert np.array_equal(after[~m], before[~m]) and not np.any(np.all(after[m] == before[m], axis=1))
    assert np.all(st1[m, -1] == 0) and np.all(st1[~m, -1] == 1)            # elapsed steps
    assert np.array_equal(st1[m, 46:48], after[m, 0:2])                      # the block sits where the new row says


@pytest.mark.parametrize("task,kw", [("reach", {}), ("pick_and_place", {}), ("block_stack", dict(num_block=4))])
def test_auto_reset_equals_step_then_reset(task, kw):
    """Two handles on the same Philox streams: one with auto-reset, one stepped and reset by hand where `done`."""
    B, T = 24, 4
    a_env = _mk(task, B, device_sampling=True, seed=9, max_episode_steps=T, check_actions=False, **kw)
    m_env = _mk(task, B, device_sampling=True, seed=9, max_episode_steps=T, check_actions=False, **kw)
    a_env.set_auto_reset(True, keep_terminal_observation=True)
    st = a_env.get_state()
    st[:, -1] = np.arange(B) % T                                           # staggered episodes
    a_env.set_state(st)
    m_env.set_state(st)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(1)
    launches0 = a_env.launch_count
    for t in range(2 * T + 1):
        a = torch.rand((B, a_env.action_dim), device="cuda", generator=gen) * 2 - 1
        oa, ra, da, ia = a_env.step(a)
        om, rm, dm, im = m_env.step(a)
        assert torch.equal(da, dm) and torch.equal(ra, rm) and torch.equal(ia["goal_achieved"], im["goal_achieved"])
        assert int(da.sum()) == B // T
        term = ia["terminal_observation"]
        for k in KEYS:
            assert torch.equal(term[k][da], om[k][da])                     # terminal rows = what the plain step returned
        om2 = m_env.reset(mask=dm)   # re-derives the rows of the envs that were NOT reset from the state (another kernel: 1e-6)
        for k in KEYS:
            assert torch.equal(oa[k][da], om2[k][dm]), (task, t, k)
            assert torch.equal(oa[k][~da], om[k][~dm]), (task, t, k)          # untouched by the auto-reset pass
            assert torch.allclose(om2[k][~dm], om[k][~dm], atol=2e-6)
        assert np.array_equal(a_env.get_state(), m_env.get_state())
    assert a_env.launch_count - launches0 == 2 * (2 * T + 1)               # step kernel + reset pass per step


def test_device_rng_is_refused_for_curriculum():
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        env = _mk("block_stack", 4, num_block=3, use_curriculum=True)
    with pytest.raises(Exception):
        env.enable_device_sampling()
