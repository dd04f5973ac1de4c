"""CPU tests of the device-side reset sampler (csrc/pmg_spawn.cuh compiled for the host by tests/emu/) against
its numpy restatement oracle/device_rng_oracle.py, and of Philox4x32-10 against the published known answers."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import device_rng_oracle as R  # noqa: E402
from tests.emu import build_emu  # noqa: E402

# Random123 kat_vectors, philox4x32 10 rounds: (counter, key) -> output
KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


@pytest.fixture(scope="module")
def emu():
    L = C.CDLL(build_emu.build())
    L.pmg_emu_device_spawn.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int64, C.c_uint32, C.POINTER(C.c_float)]
    return L


def test_philox_known_answers(emu):
    for ctr, key, want in KAT:
        assert tuple(R.philox4x32_10(ctr, key)) == want
        c, k, o = (C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), (C.c_uint32 * 4)()
        emu.pmg_emu_philox(c, k, o)
        assert tuple(o) == want


CASES = [("reach", 0, 0), ("push", 1, 0), ("pick_and_place", 1, 0), ("slide", 1, 0), ("block_stack", 4, 0), ("block_stack", 5, 1),
         ("block_stack", 2, 0), ("block_rearrange", 4, 0), ("block_rearrange", 3, 0)]


@pytest.mark.parametrize("task,nb,grip", CASES)
def test_device_spawn_rows_match_numpy_restatement(emu, task, nb, grip):
    """Bit-exact spawn rows over many (seed, env, episode) streams, and the reference's sampling rules hold."""
    tid = R.TASK_IDS[task]
    b = R.bounds(task)
    for seed, env, ep in [(0, 0, 0), (1234, 7, 3), (2 ** 40 + 5, 2 ** 33 + 1, 2 ** 31), (99, 8191, 12)] + \
                         [(5, e, p) for e in range(24) for p in range(3)]:
        want = R.sample_row(task, nb, grip, seed, env, ep)
        got = np.zeros(want.size, dtype=np.float32)
        emu.pmg_emu_device_spawn(tid, nb, grip, seed, env, ep, got.ctypes.data_as(C.POINTER(C.c_float)))
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (task, seed, env, ep, got, want)
        n = 0 if tid == 0 else (nb if tid in (3, 4) else 1)
        xy = got[:2 * n].reshape(n, 2).astype(np.float64)
        goal = got[2 * n:].astype(np.float64)
        tip = np.array(b["tip"], dtype=np.float64)
        if n:
            assert np.all(xy >= np.array(b["obj_lo"]) - 1e-6) and np.all(xy <= np.array(b["obj_hi"]) + 1e-6)
        if tid in (1, 2, 5):
            assert np.linalg.norm(xy[0] - tip[:2]) >= 0.1 - 1e-6                       # kuka_single_step_base_env.py:109
        if tid == 1:
            assert goal[2] == np.float32(0.175)
        if tid == 5:  # Slide: goals on the long table beyond the arm's reach (kuka_single_step_base_env.py:66-69)
            assert goal[2] == np.float32(0.17) and -1.09 - 1e-6 <= goal[0] <= -0.75 + 1e-6 and abs(goal[1]) <= 0.2 + 1e-6
        if tid in (3, 4):
            for i in range(n):
                assert np.linalg.norm(xy[i] - tip[:2]) > 0.06 - 1e-6
                for j in range(i):
                    assert np.linalg.norm(xy[i] - xy[j]) > 0.06 - 1e-6                 # kuka_multi_step_base_env.py:228-235
        if tid == 3:
            g = goal[:3 * n].reshape(n, 3)
            assert np.all(g[:, 0] == g[0, 0]) and np.all(g[:, 1] == g[0, 1])
            assert sorted(np.round((g[:, 2] - 0.175) / 0.03).astype(int)) == list(range(n))   # one block per level
            for i in range(n):
                assert np.linalg.norm(g[0, :2] - xy[i]) > 0.08 - 1e-6                  # kuka_multi_step_envs.py:45-53


def test_streams_are_distinct_per_env_and_episode():
    rows = {(e, p): tuple(R.sample_row("push", 1, 0, 42, e, p)) for e in range(16) for p in range(4)}
    assert len(set(rows.values())) == len(rows)


def test_pick_and_place_goal_height_rule_statistics():
    """half of the goals are on the table (kuka_single_step_base_env.py:140-143)"""
    z = np.array([R.sample_row("pick_and_place", 1, 0, 3, e, 0)[-1] for e in range(400)])
    frac = float(np.mean(z == np.float32(0.175)))
    assert 0.40 < frac < 0.60
