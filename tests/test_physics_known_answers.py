"""Physical known-answer tests of the restated contact / dynamics pipeline (oracle on CPU; the GPU
version of the same checks is marked gpu).  With no pybullet to produce goldens, closed-form
mechanics is the independent reference: free fall, Coulomb sliding at mu = 0.1 on the table, a
resting block staying put, the arm holding its pose under gravity."""
import numpy as np
import pytest

G = 9.81
DT = 0.002


def _block_state(oracle, task="push"):
    e = oracle.OracleEnv(task, binary_reward=False, max_episode_steps=10 ** 6)
    e.reset()
    s = e.get_state()
    # reset leaves the arm motors off until the first apply_action (robot_bases.py:236-238); switch them
    # on at the current pose so the arm holds still while the block is observed
    s[28:35] = s[0:7]
    s[37:44] = 200.0 * 0.04
    return e, s


def test_free_fall_matches_closed_form(oracle):
    e, s = _block_state(oracle)
    s[46:49] = [-0.52, 0.25, 0.60]       # above the table edge region, nothing underneath but air for 0.2 s
    s[53:59] = 0.0
    e.set_state(s)
    n = 100
    e.substeps(n)
    s1 = e.get_state()
    t = n * DT
    # semi-implicit Euler: v_k = -g k dt (minus the 0.04 linear damping), z = z0 - g dt^2 k(k+1)/2
    z_expected = 0.60 - G * DT * DT * n * (n + 1) / 2
    drop = 0.60 - z_expected
    # Bullet's link damping -m v (0.04 + 0.04 |v|) slows the fall by ~0.5 % over 0.2 s, never speeds it up
    assert 0.0 <= s1[48] - z_expected < 0.01 * drop
    assert abs(s1[55] + G * t) < 0.02
    assert np.allclose(s1[46:48], [-0.52, 0.25], atol=1e-12)


def test_coulomb_sliding_deceleration(oracle):
    """Block sliding on the table: deceleration mu*g with mu = 1.0 * 0.1 (block x table friction)."""
    e, s = _block_state(oracle)
    s[46:49] = [-0.60, 0.25, 0.175]
    s[53:56] = [0.5, 0.0, 0.0]           # 0.5 m/s along +x
    s[56:59] = 0.0
    # park the arm far above so it cannot interfere
    e.set_state(s)
    n = 100
    e.substeps(n)
    s1 = e.get_state()
    v_expected = 0.5 - 0.1 * G * n * DT
    assert abs(s1[53] - v_expected) < 0.02, (s1[53], v_expected)
    x_expected = -0.60 + 0.5 * n * DT - 0.5 * 0.1 * G * (n * DT) ** 2
    assert abs(s1[46] - x_expected) < 3e-3
    assert abs(s1[48] - 0.175) < 2e-4 and abs(s1[47] - 0.25) < 1e-3
    # it comes to rest and stays at rest (static friction)
    e.substeps(400)
    s2 = e.get_state()
    assert np.linalg.norm(s2[53:56]) < 2e-3
    x_stop = -0.60 + 0.5 ** 2 / (2 * 0.1 * G)
    assert abs(s2[46] - x_stop) < 0.01


def test_resting_block_stays_put(oracle):
    e, s = _block_state(oracle)
    p0 = s[46:49].copy()
    e.set_state(s)
    e.substeps(500)
    s1 = e.get_state()
    assert np.abs(s1[46:49] - p0).max() < 2e-4
    assert np.linalg.norm(s1[53:56]) < 1e-3
    assert len(e.contacts()) == 4


def test_arm_holds_pose_under_gravity(oracle):
    e = oracle.OracleEnv("reach")
    o = e.reset()
    tip0 = o["achieved_goal"].copy()
    for _ in range(5):
        o, r, d, info = e.step(np.zeros(3))
    # position motors (kp 0.03, max impulse 8 per substep) hold the 15 kg arm to well under a millimetre
    assert np.abs(o["achieved_goal"] - tip0).max() < 5e-4


def test_stack_of_two_blocks_is_stable(oracle):
    e = oracle.OracleEnv("block_stack", num_block=2, max_episode_steps=10 ** 6)
    e.reset()
    s = e.get_state()
    s[28:35] = s[0:7]
    s[37:44] = 200.0 * 0.04
    s[46:49] = [-0.60, 0.10, 0.175]
    s[46 + 13:49 + 13] = [-0.60, 0.10, 0.205]
    s[49:53] = [0, 0, 0, 1]
    s[49 + 13:53 + 13] = [0, 0, 0, 1]
    s[53:59] = 0
    s[53 + 13:59 + 13] = 0
    e.set_state(s)
    e.substeps(500)
    s1 = e.get_state()
    assert abs(s1[48] - 0.175) < 3e-4 and abs(s1[48 + 13] - 0.205) < 5e-4
    assert np.abs(s1[46:48] - [-0.60, 0.10]).max() < 1e-3 and np.abs(s1[46 + 13:48 + 13] - [-0.60, 0.10]).max() < 1e-3


@pytest.mark.gpu
def test_gpu_free_fall_and_sliding_match_closed_form():
    import contextlib
    import io
    import torch
    import pybullet_multigoal_gym_b200 as pmg
    with contextlib.redirect_stdout(io.StringIO()):
        env = pmg.make_env(task="push", batch=4, binary_reward=False, max_episode_steps=1000)
    env.reset()
    s = env.get_state()
    s[0, 46:49] = [-0.52, 0.25, 0.60]
    s[0, 53:59] = 0
    s[1, 46:49] = [-0.60, 0.25, 0.175]
    s[1, 53:56] = [0.5, 0.0, 0.0]
    s[1, 56:59] = 0
    env.set_state(s)
    obs, r, d, info = env.step(torch.zeros((4, 3), device="cuda"))
    s1 = env.get_state()
    n, t = 100, 0.2
    z_expected = 0.60 - G * DT * DT * n * (n + 1) / 2
    assert 0.0 <= s1[0, 48] - z_expected < 0.01 * (0.60 - z_expected)
    assert abs(s1[1, 53] - (0.5 - 0.1 * G * t)) < 0.02
    assert abs(s1[1, 46] - (-0.60 + 0.5 * t - 0.5 * 0.1 * G * t * t)) < 3e-3
