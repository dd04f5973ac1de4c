"""GPU test of the narrowphase fast path: box_box over a static axis-aligned box (the table, the floor) must return
bit-identical contacts to the general 15-axis search, on the device, where the step kernels run it."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rot(rng, tilt):
    yaw, ang, phi = rng.uniform(-np.pi, np.pi), rng.uniform(0, tilt), rng.uniform(-np.pi, np.pi)
    cz, sz = np.cos(yaw), np.sin(yaw)
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1.0]])
    ax = np.array([np.cos(phi), np.sin(phi), 0.0])
    K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    return (np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K) @ Rz


@pytest.mark.parametrize("stat", [1, 2])
def test_box_box_static_fast_path_is_bit_identical_on_the_device(stat):
    from pybullet_multigoal_gym_b200 import _lib
    L = _lib.load()
    rng = np.random.RandomState(7 + stat)
    table_half = np.array([0.25, 0.35, 0.08], np.float32)  # include/pmg_model_constants.h PMG_TABLE_HALF
    n = 20000
    recs = np.zeros((n, 30), np.float32)
    interior = np.zeros(n, bool)
    for i in range(n):
        half = (np.array([0.015, 0.015, 0.015]) if i % 2 else np.array([0.0125, 0.005, 0.04])).astype(np.float32)
        R = _rot(rng, [0.0, 1e-4, 0.05, 0.8, np.pi][i % 5]).astype(np.float32)
        reach = float(np.abs(R[2]) @ half)
        edge = i % 7 == 0
        x = rng.uniform(0.2, 0.3) if edge else rng.uniform(-0.15, 0.15)
        z = reach + rng.choice([-3e-2, -2e-3, -1e-4, -1e-6, 0.0, 1e-6, 1e-3])
        pS = np.array([0, 0, -table_half[2]], np.float32)   # relative to the top-face centre, as collide_pair passes them
        pD = np.array([x, rng.uniform(-0.25, 0.25), z], np.float32)
        I = np.eye(3, dtype=np.float32)
        recs[i] = np.concatenate([pS, I.ravel(), table_half, pD, R.ravel(), half] if stat == 1 else [pD, R.ravel(), half, pS, I.ravel(), table_half])
        interior[i] = not edge
    out0, out1 = np.zeros((n, 32), np.float32), np.zeros((n, 32), np.float32)
    fp = lambda a: a.ctypes.data_as(C.c_void_p)
    _lib.check(L.pmg_debug_box_box(fp(recs), n, 0, fp(out0), 0))
    _lib.check(L.pmg_debug_box_box(fp(recs), n, stat, fp(out1), 0))
    assert out0.tobytes() == out1.tobytes(), np.nonzero((out0 != out1).any(axis=1))[0][:10]
    touching = out0[:, 0] > 0
    assert touching[interior].sum() > 5000 and (out0[:, 0] == 4).sum() > 2000
    assert np.all(out0[touching, 7] <= 0)  # signed distance of the first contact
