"""CPU tests of the oracle: known answers derived from the reference's constants,
the numpy-compatible RNG (bit-exact against numpy itself), and the golden vectors produced by the
reference's own Python plumbing (tools/gen_golden.py)."""
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
KEYS = ("observation", "policy_state", "achieved_goal", "desired_goal")


def test_fk_known_answers(oracle):
    # FK(q = 0): tip straight up at 1.381 m, identity orientation (urdf joint origins summed)
    pos, quat = oracle.fk_tip(np.zeros(9))
    np.testing.assert_allclose(pos, [0, 0, 1.381], atol=1e-9)
    np.testing.assert_allclose(quat, [0, 0, 0, 1], atol=1e-9)
    # FK(kuka_rest_pose, kuka.py:27) ~ the intended start (-0.52, 0, 0.25) / quat (0,-1,0,0) up to sign
    pos, quat = oracle.fk_tip([0, -0.5592432, 0, 1.733180, 0, -0.8501557, 0, 0, 0])
    np.testing.assert_allclose(pos, [-0.52292, 0, 0.25077], atol=5e-5)
    np.testing.assert_allclose(np.abs(quat), [0, 1, 0, 0.0005], atol=5e-4)
    # kuka_away_pose, kuka.py:28
    pos, _ = oracle.fk_tip([0, 0.5467089, 0, 4.518901, 0, 0.828478, 0, 0, 0])
    np.testing.assert_allclose(pos, [0.51411, 0, 0.24801], atol=5e-5)


def test_jaw_geometry(oracle):
    # finger_closeness = 0.07 - q1 - q2 (tabs on the inner faces, urdf:485-494)
    e = oracle.OracleEnv("pick_and_place")
    e.reset()
    for q1, q2 in [(0.035, 0.035), (0.02, 0.02), (0.0, 0.0), (0.01, 0.03)]:
        s = e.get_state()
        s[7], s[8] = q1, q2
        e.set_state(s)
        d = np.linalg.norm(e.link_state(2)[:3] - e.link_state(3)[:3])
        assert abs(d - (0.07 - q1 - q2)) < 1e-12


def test_rng_matches_numpy_bit_exact(oracle):
    for seed in (0, 1, 12345, 2 ** 40 + 17):
        e = oracle.OracleEnv("reach", seed=seed)
        rs = np.random.RandomState()
        rs.seed(oracle.gym_seed_key(seed))
        assert np.array_equal(e.rng_uniform(-0.64, -0.40, 50), rs.uniform(-0.64, -0.40, 50))
        for n in (2, 4, 5, 17):
            a = np.arange(n)
            rs.shuffle(a)
            assert np.array_equal(e.rng_shuffle(n), a)
        assert np.array_equal(e.rng_uniform(0, 1, 7), rs.uniform(0, 1, 7))


def test_ik_converges_and_is_bounded(oracle):
    q0 = np.array([0, -0.5592432, 0, 1.733180, 0, -0.8501557, 0, 0.035, 0.035])
    q = oracle.ik(q0, [-0.52, 0.0, 0.25])
    pos, quat = oracle.fk_tip(q)
    assert np.linalg.norm(pos - [-0.52, 0, 0.25]) < 2e-4
    assert np.all(np.abs(q[7:] - 0.035) < 1e-15)  # finger columns of the Jacobian are zero
    # a far target cannot be reached in 40 damped iterations; per-iteration step is capped at 45 deg
    q1 = oracle.ik(q0, [-0.52, 0.0, 0.25], max_iter=1)
    assert np.max(np.abs(q1 - q0)) <= np.pi / 4 + 1e-12


def test_mass_matrix_inverse_is_spd(oracle):
    e = oracle.OracleEnv("reach")
    e.reset()
    Mi = e.minv()
    np.testing.assert_allclose(Mi, Mi.T, atol=1e-9)
    assert np.all(np.linalg.eigvalsh(Mi) > 0)
    # the two prismatic fingers carry the finger mass only along their axis: M^-1 >= 1/m
    assert Mi[7, 7] > 1.0 / 0.636951 - 1e-9


def test_reward_truth_table(oracle):
    ag = np.array([[0.0, 0.0, 0.0], [0.0, 0.0, 0.0], [0.0, 0.0, 0.0]])
    dg = np.array([[0.0, 0.0, 0.049], [0.0, 0.0, 0.05], [0.0, 0.0, 0.051]])
    r, ok = oracle.compute_reward(ag, dg, 0.05, True)
    assert list(ok) == [True, True, False]  # d > thr is the failure condition, d == thr succeeds
    assert list(r) == [0.0, 0.0, -1.0] and np.signbit(r[0])  # success is -0.0 like the reference
    r, ok = oracle.compute_reward(ag, dg, 0.05, False)
    np.testing.assert_allclose(r, [-0.049, -0.05, -0.051])
    # arbitrary leading axes (HER relabelling), and the 12-D block-stack distance
    r, ok = oracle.compute_reward(np.zeros((2, 5, 12)), np.full((2, 5, 12), 0.01), 0.05, True)
    assert r.shape == (2, 5) and ok.all()


@pytest.mark.parametrize("task,dims,adim", [("reach", [3, 3, 3, 3], 3), ("push", [20, 7, 3, 3], 3),
                                            ("pick_and_place", [20, 7, 3, 3], 4), ("block_stack", [72, 16, 12, 12], 4)])
def test_dims_and_sampling_bounds(oracle, task, dims, adim):
    e = oracle.OracleEnv(task, num_block=4, seed=3)
    assert e.dims == dims and e.adim == adim
    for _ in range(20):
        o = e.reset()
        dg = o["desired_goal"].reshape(-1, 3)
        assert np.all(dg[:, 0] >= -0.64 - 1e-12) and np.all(dg[:, 0] <= -0.40 + 1e-12)
        assert np.all(np.abs(dg[:, 1]) <= 0.15 + 1e-12)
        if task == "push":
            assert np.all(dg[:, 2] == 0.175)
        if task == "block_stack":
            assert sorted(np.round((dg[:, 2] - 0.175) / 0.03).astype(int)) == [0, 1, 2, 3]
            assert np.ptp(dg[:, 0]) == 0 and np.ptp(dg[:, 1]) == 0
            blocks = o["achieved_goal"].reshape(-1, 3)
            for i in range(4):
                for j in range(i + 1, 4):
                    assert np.linalg.norm(blocks[i, :2] - blocks[j, :2]) > 0.06
        if task != "reach":
            blocks = o["achieved_goal"].reshape(-1, 3)
            assert np.all(blocks[:, 2] == 0.175)


def test_time_limit_and_elapsed(oracle):
    e = oracle.OracleEnv("reach", max_episode_steps=5)
    e.reset()
    flags = [e.step(np.zeros(3))[2] for _ in range(6)]
    assert flags == [False, False, False, False, True, True]
    e.reset()
    assert e.step(np.zeros(3))[2] is False


def test_block_rests_on_table_and_arm_tracks(oracle):
    e = oracle.OracleEnv("push", binary_reward=False)
    o = e.reset()
    z0 = o["achieved_goal"][2]
    for _ in range(10):
        o, r, d, info = e.step(np.zeros(3))
    assert abs(o["achieved_goal"][2] - z0) < 1e-4          # resting contact holds the block
    assert len(e.contacts()) >= 4                           # four corner points in the table manifold
    e = oracle.OracleEnv("reach")
    o = e.reset()
    start = o["achieved_goal"].copy()
    for _ in range(10):
        o, r, d, info = e.step(np.array([1.0, 0.0, 0.0]))
    # target moved 10 cm in +x; the motors close ~95 % of each IK step per env-step
    assert 0.09 < o["achieved_goal"][0] - start[0] < 0.1001


# make_env / OracleEnv keyword arguments of every golden (kept in step with tools/gen_golden.py VARIANTS)
GOLDEN_VARIANTS = {
    "reach": dict(task="reach"), "push": dict(task="push", binary_reward=False),
    "pick_and_place": dict(task="pick_and_place"), "block_stack": dict(task="block_stack", num_block=4),
    "slide": dict(task="slide", binary_reward=False),
    "block_rearrange": dict(task="block_rearrange", num_block=3),
    "block_stack_grip": dict(task="block_stack", num_block=3, grip_informed_goal=True),
    "reach_jc": dict(task="reach", joint_control=True),
    "pick_and_place_jc": dict(task="pick_and_place", binary_reward=False, joint_control=True),
    "block_stack_td": dict(task="block_stack", num_block=3, task_decomposition=True),
    "block_stack_td_grip": dict(task="block_stack", num_block=3, task_decomposition=True, grip_informed_goal=True),
    "block_stack_cur": dict(task="block_stack", num_block=3, use_curriculum=True, num_goals_to_generate=12),
    "block_stack_cur_grip": dict(task="block_stack", num_block=3, use_curriculum=True, grip_informed_goal=True, num_goals_to_generate=12),
    "block_rearrange_cur": dict(task="block_rearrange", num_block=3, use_curriculum=True, num_goals_to_generate=12),
}


@pytest.mark.parametrize("name", sorted(GOLDEN_VARIANTS))
def test_oracle_reproduces_reference_plumbing_goldens(oracle, name):
    """tests/golden/ref_plumbing_*.npz were produced by the reference's unmodified Python running on
    the pybullet shim; the oracle's own C restatement of reset/step/obs/reward must reproduce them."""
    g = np.load(os.path.join(GOLDEN, "ref_plumbing_%s.npz" % name))
    e = oracle.OracleEnv(seed=0, max_episode_steps=int(g["max_episode_steps"]), **GOLDEN_VARIANTS[name])
    e.reset()  # the reference ctor's own reset (base_env.py:84)
    if "_cur" in name:
        e.set_curriculum_update(True)  # env.activate_curriculum_update() (kuka_multi_step_base_env.py:147-151)
    L = int(g["episode_len"])
    k = 0
    for ep in range(g["reset_obs"].shape[0]):
        o = e.reset()
        flat = np.concatenate([o[key] for key in KEYS])
        np.testing.assert_allclose(flat, g["reset_obs"][ep], atol=1e-12, rtol=0)
        if "_cur" in name:  # level drawn with np_random.choice, probabilities after _update_curriculum_prob
            prob, level = e.curriculum()
            assert level == int(g["curriculum_level"][ep])
            np.testing.assert_allclose(prob, g["curriculum_prob"][ep], atol=0, rtol=0)
        for t in range(L):
            for (at, ind), want in zip(g["sub_goal_calls"], g["sub_goal_returns"]):
                if at == k:  # env.set_sub_goal(ind) was called before this step (kuka_multi_step_base_env.py:159-165)
                    e.set_sub_goal(int(ind))
                    np.testing.assert_allclose(e.observe()["desired_goal"], want, atol=1e-9, rtol=0)
            o, r, done, info = e.step(g["actions"][k])
            flat = np.concatenate([o[key] for key in KEYS])
            np.testing.assert_allclose(flat, g["step_obs"][k], atol=1e-9, rtol=0, err_msg="%s step %d" % (name, k))
            assert r == g["reward"][k] and done == bool(g["done"][k]) and info["goal_achieved"] == bool(g["goal_achieved"][k])
            k += 1


def test_slide_scene_known_answers(oracle):
    """Slide (kuka_single_step_envs.py:49-59, kuka_single_step_base_env.py:53-56,66-69; long_table.urdf, cylinder_bulk.urdf):
    bounds and dimensions from the reference's constants, a puck that rests quietly on the long table, and one that
    keeps sliding on the 0.05-friction surface after the jaws hit it."""
    e = oracle.OracleEnv("slide", seed=3, binary_reward=False)
    assert e.dims == [20, 7, 3, 3] and e.adim == 3
    for _ in range(20):
        o = e.reset()
        dg, puck, tip = o["desired_goal"], o["achieved_goal"], o["observation"][:3]
        assert -1.09 - 1e-12 <= dg[0] <= -0.75 + 1e-12 and abs(dg[1]) <= 0.2 + 1e-12 and dg[2] == 0.17      # beyond the arm's reach (x >= -0.67)
        assert -0.59 - 1e-12 <= puck[0] <= -0.45 + 1e-12 and abs(puck[1]) <= 0.1 + 1e-12 and puck[2] == 0.17
        assert np.linalg.norm(puck[:2] - [-0.52, 0.0]) >= 0.1 and abs(tip[2] - 0.176) < 2e-3
    for _ in range(5):
        o, r, d, info = e.step(np.zeros(3))
    st = e.get_state()
    assert abs(st[48] - 0.17) < 2e-5 and np.abs(st[53:59]).max() < 1e-8 and len(e.contacts()) == 4   # at rest on four rim points
    assert r == -np.linalg.norm(o["achieved_goal"] - o["desired_goal"])
    # hit the puck from the +x side: it slides on towards -x after the jaws have stopped at the workspace limit
    e = oracle.OracleEnv("slide", seed=0, binary_reward=False)
    e.reset()
    x0 = e.get_state()[46]
    for t in range(34):
        st = e.get_state()
        tip = e.link_state(0)[:3]
        a = np.clip((st[46:49] + [0.06, 0.0, 0.0] - tip) / 0.01, -1, 1) if t < 12 else np.array([-1.0, 0.0, 0.0])
        e.step(a)
    st = e.get_state()
    assert st[46] < x0 - 0.1 and abs(st[48] - 0.17) < 1e-3          # pushed a long way, still flat on the table
    q = st[49:53]
    assert abs(q[0]) < 1e-3 and abs(q[1]) < 1e-3                      # no tilt (it may spin about z)
