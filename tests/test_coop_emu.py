"""CPU tests of the lane-cooperative step algorithm (csrc/pmg_coop.cuh).

The device source is compiled for the host with the 8 lanes of an environment emulated as lockstep
coroutines (tests/emu/), and compared with the double-precision CPU oracle: the joint-space inertia
inverse (lane-parallel composite-rigid-body sums + Gauss-Jordan sweeps vs the oracle's articulated-body
impulse responses), free rollouts, and rollouts that press the closed jaws onto the table (finger-table
manifolds, contact and friction rows).  The same kernel is checked on the GPU by tests/test_gpu_parity.py.
"""
import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pmg_oracle as O  # noqa: E402
from tests.emu import build_emu  # noqa: E402

FP = C.POINTER(C.c_float)
U8 = C.POINTER(C.c_uint8)


def _f(a):
    return a.ctypes.data_as(FP)


@pytest.fixture(scope="module")
def emu():
    O.build()
    L = C.CDLL(build_emu.build())
    assert L.pmg_emu_state_words() == 50
    return L


def test_shared_memory_budget(emu):
    # 4 environments per one-warp block, 14 blocks per SM (batch 8192 on 148 SMs) must fit in 227 KB
    per_block = emu.pmg_emu_table_bytes() + 4 * emu.pmg_emu_smem_bytes() + 1024
    assert 14 * per_block <= 227 * 1024


@pytest.mark.parametrize("seed", [0, 3])
def test_mass_matrix_inverse_matches_oracle(emu, seed):
    o = O.OracleEnv("reach", seed=seed)
    o.reset()
    rng = np.random.RandomState(seed)
    st = o.get_state()
    st[:7] += rng.uniform(-0.3, 0.3, 7)
    st[7:9] = rng.uniform(0.0, 0.035, 2)
    o.set_state(st)
    q = st[:9].astype(np.float32)
    qd = np.zeros(9, np.float32)
    minv, qo, qdo = np.zeros(81, np.float32), np.zeros(9, np.float32), np.zeros(9, np.float32)
    assert emu.pmg_emu_substep(_f(q), _f(qd), _f(minv), _f(qo), _f(qdo)) == 0
    ref = o.minv()
    got = minv.reshape(9, 9)
    assert np.abs(got - ref).max() <= 2e-6 * np.abs(ref).max()
    assert np.array_equal(got, got.T) or np.abs(got - got.T).max() <= 1e-6 * np.abs(ref).max()


def _rollout(emu, seed, nsteps, press_down):
    o = O.OracleEnv("reach", seed=seed)
    o.reset()
    o.reset()
    st = o.get_state().astype(np.float32)
    man = np.zeros(82, np.float32)
    rng = np.random.RandomState(seed + 100)
    obs, rew = np.zeros(12, np.float32), np.zeros(1, np.float32)
    dn, su = np.zeros(1, np.uint8), np.zeros(1, np.uint8)
    worst, points = 0.0, 0
    for t in range(nsteps):
        a = rng.uniform(-1, 1, 3).astype(np.float32)
        if press_down:
            a[2] = -1.0
        ro, rr, rd, ri = o.step(a.astype(np.float64))
        rc = emu.pmg_emu_reach_step(_f(st), _f(man), _f(a), C.c_float(0.05), 1, 50, _f(obs), _f(rew),
                                    dn.ctypes.data_as(U8), su.ctypes.data_as(U8))
        assert rc == 0, "divergent collective in the cooperative kernel"
        worst = max(worst, float(np.abs(obs[6:9] - ro["achieved_goal"]).max()))
        assert np.array_equal(obs[0:3], obs[6:9]) and np.array_equal(obs[3:6], obs[6:9])
        assert np.allclose(obs[9:12], ro["desired_goal"], atol=1e-7)
        assert bool(dn[0]) == rd and bool(su[0]) == ri["goal_achieved"] and rew[0] == np.float32(rr)
        points = max(points, int(man[0:1].view(np.int32)[0]) + int(man[41:42].view(np.int32)[0]))
    return worst, points, st, o.get_state()


def test_free_rollout_matches_oracle(emu):
    worst, points, st, ost = _rollout(emu, seed=1, nsteps=12, press_down=False)
    assert worst < 1e-5, worst  # tolerance of the path is 1e-4 (BASELINE.json north_star)
    assert np.abs(st[:18] - ost[:18]).max() < 1e-4
    assert st[49] == 12.0


def test_finger_table_contact_rollout_matches_oracle(emu):
    # tip starts at z = 0.25; 8+ steps of -1 cm reach the 0.175 clip where the closed jaws touch the table
    worst, points, st, ost = _rollout(emu, seed=2, nsteps=14, press_down=True)
    assert points == 8, points  # both finger-table manifolds full: 4 + 4 points
    assert worst < 1e-4, worst


@pytest.mark.parametrize("task,nb", [("push", 1), ("block_stack", 3)])
def test_thread_per_env_physics_on_host_matches_oracle(emu, task, nb):
    """The thread-per-env device code (csrc/pmg_sim.cuh: robot dynamics, box-box manifolds, contact / friction
    rows, PGS) compiled for the host, stepped from the oracle's state with blocks resting on the table and the
    closed jaws descending onto it: positions agree with the double-precision oracle to 1e-5 over 0.2 s, velocities
    to 1e-4 (contact depths are evaluated relative to the reference face / the static box's anchor, so a resting
    block does not pick up fp32 rounding of table-sized coordinates as spin: 3e-4 rad/s before that change)."""
    o = O.OracleEnv(task, num_block=nb, seed=4)
    o.reset()
    o.reset()
    a = np.zeros(3 if task == "push" else 4)
    a[2] = -1.0
    o.step(a)   # motors on, arm moving down
    st = o.get_state().astype(np.float32)
    o.set_state(st.astype(np.float64))
    npairs = 2 + 4 * nb + nb * (nb - 1) // 2 + (nb if nb >= 2 else 0)  # + gripper base against every block
    man = np.zeros(41 * npairs, np.float32)
    s2 = st.copy()
    for call in range(5):
        assert emu.pmg_emu_thread_substeps(nb, _f(s2), _f(man), 1) == 0
        o.step_simulation()
        ref = o.get_state()
        assert np.abs(s2[:9] - ref[:9]).max() < 1e-5                     # joint angles
        for b in range(nb):
            got, want = s2[46 + 13 * b:59 + 13 * b], ref[46 + 13 * b:59 + 13 * b]
            assert np.abs(got[:7] - want[:7]).max() < 1e-5               # block position + quaternion
            assert np.abs(got[7:10] - want[7:10]).max() < 1e-4           # linear velocity
            assert np.abs(got[10:13] - want[10:13]).max() < 1e-4         # angular velocity
    assert sum(int(man[41 * k:41 * k + 1].view(np.int32)[0]) for k in range(npairs)) >= 4 * nb  # blocks rest on 4 points


@pytest.mark.parametrize("task,tid,adim,nsteps", [("push", 1, 3, 16), ("pick_and_place", 2, 4, 30)])
def test_cooperative_block_step_matches_oracle(emu, task, tid, adim, nsteps):
    """The lane-cooperative Push / PickAndPlace step (block + six manifolds in shared memory, rows with a block end
    point, lanes over pairs for the narrowphase and over contact points for the row set-up) against the oracle,
    one env.step at a time from the oracle's fp32-rounded state (contact-rich rollouts are chaotic in open loop):
    scripted side push / descend-and-grasp, every position entry of the packed row within 1e-4."""
    assert 7 * (emu.pmg_emu_table_bytes() + 4 * emu.pmg_emu_block_smem_bytes() + 1024) <= 227 * 1024  # 7 blocks per SM
    o = O.OracleEnv(task, seed=1, binary_reward=False)
    o.reset()
    o.reset()
    obs, rew = np.zeros(33, np.float32), np.zeros(1, np.float32)
    dn, su = np.zeros(1, np.uint8), np.zeros(1, np.uint8)
    worst, touched, most = 0.0, 0, 0
    for t in range(nsteps):
        st = o.get_state().astype(np.float32)
        o.set_state(st.astype(np.float64))     # both sides start the step from the same state, contact caches empty
        man = np.zeros(6 * 41, np.float32)
        tip = o.link_state(0)[:3]
        a = np.zeros(adim, np.float32)
        # push: straight at the block; pick and place: hover, descend, close the jaws (t >= 16), lift (t >= 22)
        dz = 0.0 if task == "push" else (0.07 if t <= 10 else (0.0 if t < 22 else 0.08))
        a[:3] = np.clip((st[46:49] + np.array([0.0, 0.0, dz]) - tip) / 0.01, -1, 1)
        if adim == 4:
            a[3] = -1.0 if t < 16 else 1.0
        ro, rr, rd, ri = o.step(a.astype(np.float64))
        rc = emu.pmg_emu_block_step(tid, _f(st), _f(man), _f(a), C.c_float(0.05), 0, 50, _f(obs), _f(rew),
                                    dn.ctypes.data_as(U8), su.ctypes.data_as(U8))
        assert rc == 0, "divergent collective in the cooperative kernel"
        want = np.concatenate([ro[k] for k in ("observation", "policy_state", "achieved_goal", "desired_goal")])
        pos = np.r_[0:10, 20:33]               # everything but the velocity entries of the observation
        worst = max(worst, float(np.abs(obs - want)[pos].max()))
        assert abs(float(rew[0]) - rr) < 1e-4 and bool(dn[0]) == rd
        touched = max(touched, sum(int(man[41 * k:41 * k + 1].view(np.int32)[0]) for k in (4, 5)))
        most = max(most, sum(int(man[41 * k:41 * k + 1].view(np.int32)[0]) for k in range(6)))
    assert worst < 1e-4, worst
    assert touched > 0                         # the jaws did reach the block (finger-block manifolds in use)
    print("%s: most cached contact points in one step: %d" % (task, most))
    if task == "pick_and_place":
        print("pick_and_place: block height after the lift %.4f (oracle %.4f)" % (obs[5], ro["observation"][5]))
        assert obs[5] > 0.19 and ro["observation"][5] > 0.19   # carried by the friction rows with both end points
    if task == "push":
        assert most > 12                       # rows beyond the 12 shared-memory points: the global spill path ran


@pytest.mark.parametrize("task,tid,adim,width", [("reach", 0, 7, 26), ("push", 1, 7, 47), ("pick_and_place", 2, 8, 47)])
def test_cooperative_joint_control_step_matches_oracle(emu, task, tid, adim, width):
    """joint_control=True on the cooperative kernels (kuka.py:204-206: motor targets += 0.05 a, no IK; the 7 joint
    positions prepended to observation and policy_state), one env.step at a time from the oracle's state."""
    o = O.OracleEnv(task, seed=3, binary_reward=False, joint_control=True)
    o.reset()
    o.reset()
    rng = np.random.RandomState(11)
    obs, rew = np.zeros(width, np.float32), np.zeros(1, np.float32)
    dn, su = np.zeros(1, np.uint8), np.zeros(1, np.uint8)
    nman = 2 if tid == 0 else 6
    worst = 0.0
    for t in range(6):
        st = o.get_state().astype(np.float32)
        o.set_state(st.astype(np.float64))
        man = np.zeros(nman * 41, np.float32)
        a = rng.uniform(-1, 1, adim).astype(np.float32)
        ro, rr, rd, ri = o.step(a.astype(np.float64))
        rc = emu.pmg_emu_step_jc(tid, _f(st), _f(man), _f(a), C.c_float(0.05), 0, 50, _f(obs), _f(rew),
                                 dn.ctypes.data_as(U8), su.ctypes.data_as(U8))
        assert rc == 0, "divergent collective in the cooperative kernel"
        want = np.concatenate([ro[k] for k in ("observation", "policy_state", "achieved_goal", "desired_goal")])
        assert want.shape == (width,)
        # everything but the velocity entries of the (prefixed) block observation
        pos = np.arange(width) if tid == 0 else np.r_[0:17, 27:47]
        worst = max(worst, float(np.abs(obs - want)[pos].max()))
        assert abs(float(rew[0]) - rr) < 1e-4 and bool(dn[0]) == rd
    assert worst < 1e-4, worst


def _block_scenario(emu, task, tid, adim, nsteps, init, policy):
    """Teacher-forced cooperative block steps against the oracle; returns (worst position-entry error, set of
    collision pairs that held contact points at a step end, most points at a step end)."""
    o = O.OracleEnv(task, seed=4, binary_reward=False)
    o.reset()
    o.reset()
    st = o.get_state()
    init(st)
    o.set_state(st)
    obs, rew = np.zeros(33, np.float32), np.zeros(1, np.float32)
    dn, su = np.zeros(1, np.uint8), np.zeros(1, np.uint8)
    worst, most, pairs = 0.0, 0, set()
    pos = np.r_[0:10, 20:33]
    for t in range(nsteps):
        st = o.get_state().astype(np.float32)
        o.set_state(st.astype(np.float64))
        man = np.zeros(6 * 41, np.float32)
        a = policy(t, st, o.link_state(0)[:3]).astype(np.float32)
        ro, rr, rd, ri = o.step(a.astype(np.float64))
        rc = emu.pmg_emu_block_step(tid, _f(st), _f(man), _f(a), C.c_float(0.05), 0, 50, _f(obs), _f(rew),
                                    dn.ctypes.data_as(U8), su.ctypes.data_as(U8))
        assert rc == 0, "divergent collective in the cooperative kernel"
        want = np.concatenate([ro[k] for k in ("observation", "policy_state", "achieved_goal", "desired_goal")])
        worst = max(worst, float(np.abs(obs - want)[pos].max()))
        cnt = [int(man[41 * k:41 * k + 1].view(np.int32)[0]) for k in range(6)]
        most = max(most, sum(cnt))
        pairs |= {k for k in range(6) if cnt[k]}
    return worst, pairs, most, float(obs[5])


def test_cooperative_block_falls_to_the_floor_like_the_oracle(emu):
    """A block released beyond the table edge: free fall, impact on the floor box (pair 3), bounce from the
    penetration recovery, rest -- the floor-block rows of the cooperative kernel."""
    def init(st):
        st[46:49] = [-0.52, 0.37, 0.30]
        st[53:59] = 0.0
    worst, pairs, most, z = _block_scenario(emu, "push", 1, 3, 9, init, lambda t, st, tip: np.zeros(3))
    assert worst < 1e-4, worst
    assert 3 in pairs and z < 0.03   # on the floor at the end


def test_cooperative_closed_jaws_press_on_the_block_like_the_oracle(emu):
    """Hover over the block, then drive the closed jaws down onto it: finger-block rows loaded against the table-block
    rows (pairs 2, 4, 5 together)."""
    def policy(t, st, tip):
        a = np.zeros(4)
        a[:3] = np.clip((st[46:49] + np.array([0.0, 0.0, 0.08 if t <= 6 else 0.0]) - tip) / 0.01, -1, 1)
        if t > 12:
            a[2] = -1.0
        a[3] = 1.0
        return a
    worst, pairs, most, z = _block_scenario(emu, "pick_and_place", 2, 4, 16, lambda st: None, policy)
    assert worst < 1e-4, worst
    assert {2, 4, 5} <= pairs and most >= 10


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_cooperative_tumbling_block_matches_oracle(emu, seed):
    """A spinning block dropped onto the table with a random orientation: edge and corner impacts (edge-edge SAT axes,
    general clipping, manifold replacement) until it settles on a face.  One 0.2 s step at a time from the oracle's
    state; impacts amplify fp32 rounding, hence 3e-4 here (observed <= 1e-4)."""
    rng = np.random.RandomState(seed)
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w = rng.normal(size=3) * 3

    def init(st):
        st[46:49] = [-0.52, 0.1, 0.22]
        st[49:53] = q
        st[53:56] = [0.1, 0.0, 0.0]
        st[56:59] = w
    worst, pairs, most, z = _block_scenario(emu, "push", 1, 3, 8, init, lambda t, st, tip: np.zeros(3))
    assert worst < 3e-4, worst
    assert pairs == {2} and most == 4 and abs(z - 0.175) < 1e-3   # flat on the table at the end


def _multi_scenario(emu, nb, nsteps, window, init, policy, seed=4, grip=False, td=False, sub_goal=None, task="block_stack", cur=False):
    """Teacher-forced cooperative multi-block steps against the oracle (state and packed observation row); the
    emulator runs only for t in `window` (the oracle alone drives the approach).  Returns worst joint/block pose error,
    worst block velocity error, worst observation-row error away from the velocity entries, the collision pairs that
    held points at a step end, most points at a step end."""
    o = O.OracleEnv(task, num_block=nb, seed=seed, binary_reward=False, grip_informed_goal=grip, task_decomposition=td, use_curriculum=cur)
    o.reset()
    o.reset()
    if cur:  # the step kernels read the curriculum word like a sub-goal index (block_stack) / as the mask of moved blocks (block_rearrange: td = 2)
        td = 2 if task == "block_rearrange" else 1
    if sub_goal is not None:
        o.set_sub_goal(sub_goal)
    st = o.get_state()
    init(st)
    o.set_state(st)
    npairs = 2 + 4 * nb + nb * (nb - 1) // 2 + (nb if nb >= 2 else 0)  # + gripper base against every block
    G = 3 * nb + (4 if grip else 0)
    width = (8 + 16 * nb) + (4 + 3 * nb) + 2 * G
    ovf = (C.c_int * 1)(0)
    obs, rew = np.zeros(width, np.float32), np.zeros(1, np.float32)
    dn, su = np.zeros(1, np.uint8), np.zeros(1, np.uint8)
    vel_cols = set(range(4, 8))
    for b in range(nb):
        vel_cols |= set(range(8 + 16 * b + 10, 8 + 16 * b + 16))
    pos_cols = np.array([c for c in range(width) if c not in vel_cols])
    worst_p, worst_v, worst_o, pairs, most = 0.0, 0.0, 0.0, set(), 0
    for t in range(nsteps):
        st = o.get_state().astype(np.float32)
        o.set_state(st.astype(np.float64))
        a = policy(t, st, o.link_state(0)[:3]).astype(np.float32)
        ro, rr, rd, ri = o.step(a.astype(np.float64))
        if t not in window:
            continue
        man = np.zeros(41 * npairs, np.float32)
        s2 = st.copy()
        rc = emu.pmg_emu_multi_step(nb, _f(s2), _f(man), _f(a), ovf, -1 if task == "block_rearrange" else int(grip), int(td), C.c_float(0.05), 0, 50, _f(obs), _f(rew),
                                    dn.ctypes.data_as(U8), su.ctypes.data_as(U8))
        assert rc == 0, "divergent collective in the cooperative kernel"
        ref = o.get_state()
        worst_p = max(worst_p, float(np.abs(s2[:9] - ref[:9]).max()))
        for b in range(nb):
            worst_p = max(worst_p, float(np.abs(s2[46 + 13 * b:53 + 13 * b] - ref[46 + 13 * b:53 + 13 * b]).max()))
            worst_v = max(worst_v, float(np.abs(s2[53 + 13 * b:59 + 13 * b] - ref[53 + 13 * b:59 + 13 * b]).max()))
        want = np.concatenate([ro[k] for k in ("observation", "policy_state", "achieved_goal", "desired_goal")])
        assert want.shape == (width,)
        worst_o = max(worst_o, float(np.abs(obs - want)[pos_cols].max()))
        assert abs(float(rew[0]) - rr) < 2e-4 and bool(dn[0]) == rd
        assert s2[-1] == ref[-1]                                   # elapsed steps
        cnt = [int(man[41 * k:41 * k + 1].view(np.int32)[0]) for k in range(npairs)]
        pairs |= {k for k in range(npairs) if cnt[k]}
        most = max(most, sum(cnt))
    assert ovf[0] == 0
    return worst_p, worst_v, worst_o, pairs, most


def test_cooperative_multi_block_grasp_matches_oracle(emu):
    """Multi-block cooperative step (lane b owns block b, lanes stride over the 17 collision pairs, rows with up to two
    block ends, block deltas fetched by run-time-source shuffles): descend onto block 0 of three, close the jaws, start
    the lift, the other two blocks resting beside it.  Shared memory: 4 environments x 4 blocks per SM fit."""
    assert 4 * (emu.pmg_emu_table_bytes() + 4 * emu.pmg_emu_multi_smem_bytes(4) + 1024) <= 227 * 1024

    def policy(t, st, tip):
        b0 = st[46:49]
        tgt = b0 + [0, 0, 0.07] if t <= 8 else (b0 if t <= 16 else (tip if t <= 20 else np.array([tip[0], tip[1], 0.27])))
        a = np.zeros(4)
        a[:3] = np.clip((tgt - tip) / 0.01, -1, 1)
        a[3] = -1.0 if t <= 16 else 1.0
        return a
    worst_p, worst_v, worst_o, pairs, most = _multi_scenario(emu, 3, 24, range(15, 24), lambda st: None, policy)
    assert worst_p < 1e-4 and worst_v < 1e-3 and worst_o < 1e-4, (worst_p, worst_v, worst_o)
    assert {4, 5, 6, 10} <= pairs and most >= 16   # both jaws on block 0, blocks 1 and 2 on the table


def test_cooperative_multi_block_stack_matches_oracle(emu):
    """Block 0 released 2 mm above block 1 (block-block pair: rows with two block ends), then the closed jaws are
    lowered gently onto the stack (finger-block, block-block and table-block rows loaded in series)."""
    def init(st):
        st[46:49] = st[59:62] + [0.0, 0.0, 0.032]
        st[49:53] = [0, 0, 0, 1]
        st[53:59] = 0.0

    def policy(t, st, tip):
        top = st[46:49]
        tgt = top + [0, 0, 0.08] if t <= 6 else top + [0, 0, 0.035]
        a = np.zeros(4)
        a[:3] = np.clip((tgt - tip) / 0.01, -1, 1)
        if t > 12:
            a[2] = -0.4
        a[3] = 1.0
        return a
    worst_p, worst_v, worst_o, pairs, most = _multi_scenario(emu, 2, 17, [0, 1, 2, 13, 14, 15, 16], init, policy)
    assert worst_p < 1e-4 and worst_v < 1e-3 and worst_o < 1e-4, (worst_p, worst_v, worst_o)
    assert {4, 5, 6, 10} <= pairs and most >= 16   # pair 2 + 4 * 2 = 10: block 0 on block 1; both jaws on block 0


@pytest.mark.parametrize("grip,td,sub_goal", [(True, False, None), (False, True, 1), (True, True, 2)])
def test_cooperative_multi_block_goal_variants_match_oracle(emu, grip, td, sub_goal):
    """Observation assembly of the multi-block cooperative step for the grip-informed goal and the task-decomposition
    sub-goals (desired goal rebuilt from the current block positions), three resting blocks, arm moving."""
    def policy(t, st, tip):
        return np.array([0.5, -0.5, -0.3, -1.0])
    worst_p, worst_v, worst_o, pairs, most = _multi_scenario(emu, 3, 2, range(0, 2), lambda st: None, policy, grip=grip, td=td, sub_goal=sub_goal)
    assert worst_p < 1e-4 and worst_o < 1e-4, (worst_p, worst_o)


@pytest.mark.parametrize("nb", [4, 5])
def test_cooperative_multi_block_larger_scenes_match_oracle(emu, nb):
    """The 4- and 5-block instantiations (24 / 32 collision pairs: lanes run up to four pairs each; 16 / 20 resting
    contact points, i.e. spilled rows): two steps with the arm moving over resting blocks."""
    def policy(t, st, tip):
        return np.array([-0.5, 0.5, -0.5, 1.0])
    worst_p, worst_v, worst_o, pairs, most = _multi_scenario(emu, nb, 2, range(0, 2), lambda st: None, policy)
    assert worst_p < 1e-4 and worst_v < 1e-3 and worst_o < 1e-4, (worst_p, worst_v, worst_o)
    assert most == 4 * nb and pairs == {2 + 4 * b for b in range(nb)}


def test_cooperative_multi_block_pick_carry_release_matches_oracle(emu):
    """The whole stacking move on three blocks -- hover, descend, grasp, lift, carry over block 1, lower, open the jaws
    -- driven by the oracle; the cooperative step is checked while carrying (block held by friction rows only), at the
    release and with block 0 resting on block 1 (block-block pair 14)."""
    def policy(t, st, tip):
        b0, b1 = st[46:49], st[59:62]
        if t <= 8:
            tgt = b0 + [0, 0, 0.07]
        elif t <= 16:
            tgt = b0
        elif t <= 20:
            tgt = tip.copy()
        elif t <= 28:
            tgt = np.array([tip[0], tip[1], 0.27])
        elif t <= 44:
            tgt = np.array([b1[0], b1[1], 0.27])
        else:
            tgt = np.array([b1[0], b1[1], 0.175 + 0.03 + 0.004])
        a = np.zeros(4)
        a[:3] = np.clip((tgt - tip) / 0.01, -1, 1)
        a[3] = -1.0 if (t <= 16 or t > 56) else 1.0
        return a
    worst_p, worst_v, worst_o, pairs, most = _multi_scenario(emu, 3, 60, [36, 50, 56, 57, 59], lambda st: None, policy)
    assert worst_p < 1e-4 and worst_v < 1e-3 and worst_o < 1e-4, (worst_p, worst_v, worst_o)
    assert {4, 5, 14} <= pairs   # both jaws on block 0 while carrying; block 0 on block 1 at the end


def test_cooperative_multi_block_gripper_base_lands_on_a_stack(emu):
    """The gripper-base cylinder (iiwa14_parallel_jaw.urdf:399-416: r 0.05, 0.045 - 0.085 above the tip) against the
    blocks (hard part 4): with the jaws open around a three-block stack the arm comes down until
    the base sits on the top block (pair 2 + 4*3 + 3 + 2 = 19: rows with a robot end that has no finger column and a
    block end) and presses on it.  Every step from the approach over the first touch to the loaded stack against the
    oracle (the topple that follows is a bifurcation: not compared)."""
    sx, sy = -0.45, 0.10

    def init(st):
        for b in range(3):
            st[46 + 13 * b:49 + 13 * b] = [sx, sy, 0.175 + 0.03 * b]
            st[49 + 13 * b:53 + 13 * b] = [0, 0, 0, 1]
            st[53 + 13 * b:59 + 13 * b] = 0.0

    def policy(t, st, tip):
        tgt = np.array([tip[0], tip[1], 0.32]) if t < 8 else (np.array([sx, sy, 0.32]) if t < 24 else np.array([sx, sy, 0.19]))
        a = np.zeros(4)
        a[:3] = np.clip((tgt - tip) / 0.01, -1, 1)
        a[3] = -1.0
        return a
    worst_p, worst_v, worst_o, pairs, most = _multi_scenario(emu, 3, 38, range(33, 38), init, policy, seed=3)
    assert worst_p < 1e-4 and worst_v < 2e-3 and worst_o < 1e-4, (worst_p, worst_v, worst_o)
    assert {2, 14, 16, 19} <= pairs   # the stack on the table, block on block twice, the gripper base on the top block


def test_cooperative_multi_block_rearrange_matches_oracle(emu):
    """BlockRearrange runs on the same multi-block step (3 action columns, jaws kept closed, table targets instead of a
    stack as the desired goal, zero finger entries in the observation)."""
    def policy(t, st, tip):
        return np.array([0.3, 0.6, -0.4])
    worst_p, worst_v, worst_o, pairs, most = _multi_scenario(emu, 3, 2, range(0, 2), lambda st: None, policy, task="block_rearrange")
    assert worst_p < 1e-4 and worst_v < 1e-3 and worst_o < 1e-4, (worst_p, worst_v, worst_o)


@pytest.mark.parametrize("level", [0, 1])
def test_cooperative_multi_block_rearrange_curriculum_goal_matches_oracle(emu, level):
    """BlockRearrange with the curriculum (kuka_multi_step_envs.py:193-227): the blocks of the episode's mask keep their
    sampled targets, every other block's desired goal follows the block; the tip pushes block 0 along."""
    def policy(t, st, tip):
        return np.clip((st[46:49] - tip) / 0.01, -1, 1)

    def init(st):
        assert st[-2] in (1.0, 2.0, 4.0)           # level 0: one moved block
        if level == 1:
            st[-2] = 5.0                            # blocks 0 and 2 moved: their goal words are the targets
    worst_p, worst_v, worst_o, pairs, most = _multi_scenario(emu, 3, 4, range(0, 4), init, policy, task="block_rearrange", cur=True)
    assert worst_p < 1e-4 and worst_v < 1e-3 and worst_o < 1e-4, (worst_p, worst_v, worst_o)


@pytest.mark.parametrize("seed", [0, 1])
def test_cooperative_slide_step_matches_oracle(emu, seed):
    """Slide: the one-block cooperative step instantiated with the long table and the puck
    (cylinder-box narrowphase: cap rim points on the table, curved side against the jaws, a jaw face on the cap;
    anisotropic inertia through the scaled lever arm of the rows).  Teacher-forced against the oracle: go behind the
    puck, hit it towards -x, come down on top of it; position entries of the packed row within 1e-4 except on the
    oracle's own bifurcation steps (a rim point entering the jaw's face), bounded at 2e-3."""
    o = O.OracleEnv("slide", seed=seed, binary_reward=False)
    o.reset()
    o.reset()
    obs, rew = np.zeros(33, np.float32), np.zeros(1, np.float32)
    dn, su = np.zeros(1, np.uint8), np.zeros(1, np.uint8)
    errs, side, cap = [], 0, 0
    xy0 = o.get_state()[46:48].copy()
    for t in range(36):
        st = o.get_state().astype(np.float32)
        o.set_state(st.astype(np.float64))
        tip, puck = o.link_state(0)[:3], st[46:49]
        if t < 12:
            a = np.clip((puck + np.array([0.06, 0.0, 0.0]) - tip) / 0.01, -1, 1)
        elif t < 24:
            a = np.array([-1.0, 0.0, 0.0])
        else:
            a = np.clip((puck + np.array([0.0, 0.0, 0.03]) - tip) / 0.01, -1, 1)
        a = a.astype(np.float32)
        ro, rr, rd, ri = o.step(a.astype(np.float64))
        man = np.zeros(6 * 41, np.float32)
        s2 = st.copy()
        rc = emu.pmg_emu_block_step(5, _f(s2), _f(man), _f(a), C.c_float(0.05), 0, 50, _f(obs), _f(rew),
                                    dn.ctypes.data_as(U8), su.ctypes.data_as(U8))
        assert rc == 0, "divergent collective in the cooperative kernel"
        want = np.concatenate([ro[k] for k in ("observation", "policy_state", "achieved_goal", "desired_goal")])
        pos = np.r_[0:10, 20:33]
        errs.append(float(np.abs(obs - want)[pos].max()))
        assert abs(float(rew[0]) - rr) < 2e-3 and bool(dn[0]) == rd
        cnt = [int(man[41 * k:41 * k + 1].view(np.int32)[0]) for k in range(6)]
        assert cnt[2] == 4 and cnt[3] == 0           # the puck rests on the long table on its four rim points
        side += cnt[4] + cnt[5]
    errs = np.array(errs)
    assert np.mean(errs < 1e-4) >= 0.9 and errs.max() < 2e-3, errs
    assert np.linalg.norm(o.get_state()[46:48] - xy0) > 0.003 and side > 0   # the jaws touched and moved the puck


def _rot(rng, tilt):
    """Rotation about a random axis: by up to `tilt` radians about a horizontal axis after a free yaw."""
    yaw, ang, phi = rng.uniform(-np.pi, np.pi), rng.uniform(0, tilt), rng.uniform(-np.pi, np.pi)
    cz, sz = np.cos(yaw), np.sin(yaw)
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1.0]])
    ax = np.array([np.cos(phi), np.sin(phi), 0.0])
    K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    Rt = np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K
    return Rt @ Rz


@pytest.mark.parametrize("stat", [1, 2])
def test_box_box_static_fast_path_is_bit_identical(emu, stat):
    """box_box(stat = 1 / 2) -- the separating-axis search cut down to four face axes for a box lying over a static
    box's top face, well inside its outline (pmg_physics.cuh) -- returns exactly the contacts of the full 15-axis
    search: random fingers / cubes over the table, flat, tilted and on an edge, touching, deep and clear of it, and
    near and across the table's sides, where the fast path must hand over to the general one."""
    rng = np.random.RandomState(40 + stat)
    table_half = np.array([0.4, 0.5, 0.0775], np.float32)  # order of magnitude of the reference's tables
    out0, out1 = np.zeros(32, np.float32), np.zeros(32, np.float32)
    touching = fast_cases = 0
    for case in range(4000):
        half = (np.array([0.02, 0.02, 0.02]) if case % 2 else np.array([0.01, 0.004, 0.04])).astype(np.float32)
        R = _rot(rng, [0.0, 1e-4, 0.05, 0.8, np.pi][case % 5]).astype(np.float32)
        reach = float(np.abs(R[2]) @ half)  # extent along z
        edge = case % 7 == 0                # over a side of the table: general path
        x = rng.uniform(0.36, 0.46) if edge else rng.uniform(-0.3, 0.3)
        z = table_half[2] + reach + rng.choice([-3e-2, -2e-3, -1e-4, -1e-6, 0.0, 1e-6, 1e-3])
        pD = np.array([x, rng.uniform(-0.4, 0.4), z], np.float32)
        # coordinates relative to the table's top-face centre, as collide_pair passes them
        pS = np.array([0, 0, -table_half[2]], np.float32)
        pD = pD + pS
        I = np.eye(3, dtype=np.float32)
        if stat == 1:
            rec = np.concatenate([pS, I.ravel(), table_half, pD, R.ravel(), half])
        else:
            rec = np.concatenate([pD, R.ravel(), half, pS, I.ravel(), table_half])
        rec = np.ascontiguousarray(rec, np.float32)
        out0[:] = 0; out1[:] = 0
        n0 = emu.pmg_emu_box_box(_f(rec), 0, _f(out0))
        n1 = emu.pmg_emu_box_box(_f(rec), stat, _f(out1))
        assert n0 == n1 and out0.tobytes() == out1.tobytes(), (case, n0, n1, out0[:1 + 7 * n0], out1[:1 + 7 * n1])
        touching += n0 > 0
        fast_cases += (not edge)
    assert touching > 1000 and fast_cases > 3000
