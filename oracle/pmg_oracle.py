"""ctypes binding of the CPU oracle (oracle/libpmg_oracle.so).

TEST INFRASTRUCTURE, NOT PRODUCT CODE: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.  PARITY UNPINNED for the physics
(see pmg_oracle.h).
"""
import ctypes as C
import hashlib
import os
import struct
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

TASKS = {"reach": 0, "push": 1, "pick_and_place": 2, "block_stack": 3, "block_rearrange": 4, "slide": 5}


def build(force=False):
    so = os.path.join(_HERE, "libpmg_oracle.so")
    src = os.path.join(_HERE, "pmg_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"])
    return so


_NATIVE_SO = None


def use_native():
    """CPU-baseline runs (bench.py): compile the same source on THIS machine with -O3 -march=native (still
    -fno-fast-math -ffp-contract=off, so the arithmetic is unchanged) into oracle/_native/ and load that build
    instead of the portable -O2 one.  Falls back to the portable build when no compiler is available.  Returns the
    flags in use."""
    global _NATIVE_SO
    if _LIB is not None:
        return "already loaded"
    out_dir = os.path.join(_HERE, "_native")
    so = os.path.join(out_dir, "libpmg_oracle_native.so")
    flags = ["-O3", "-march=native", "-fPIC", "-std=gnu11", "-fno-fast-math", "-ffp-contract=off"]
    try:
        os.makedirs(out_dir, exist_ok=True)
        subprocess.check_call(["gcc"] + flags + ["-shared", "-o", so, os.path.join(_HERE, "pmg_oracle.c"), "-lm", "-lpthread"],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        _NATIVE_SO = so
        return " ".join(flags)
    except Exception:
        build()
        return "-O2 (portable build; native compile failed)"


def lib():
    global _LIB
    if _LIB is None:
        so = _NATIVE_SO or os.path.join(_HERE, "libpmg_oracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
        L.pmgo_create.restype = C.c_void_p
        L.pmgo_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_int]
        L.pmgo_create_ex.restype = C.c_void_p
        L.pmgo_create_ex.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int]
        L.pmgo_create_ex2.restype = C.c_void_p
        L.pmgo_create_ex2.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int]
        L.pmgo_create_ex3.restype = C.c_void_p
        L.pmgo_create_ex3.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_long]
        L.pmgo_set_curriculum_update.argtypes = [C.c_void_p, C.c_int]
        L.pmgo_get_curriculum.argtypes = [C.c_void_p, dp]
        L.pmgo_get_curriculum.restype = C.c_int
        L.pmgo_set_sub_goal.argtypes = [C.c_void_p, C.c_int]
        L.pmgo_observe.argtypes = [C.c_void_p, dp]
        L.pmgo_destroy.argtypes = [C.c_void_p]
        L.pmgo_dims.argtypes = [C.c_void_p, ip]
        L.pmgo_dims.restype = C.c_int
        L.pmgo_seed_array.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.c_int]
        L.pmgo_reset.argtypes = [C.c_void_p, dp]
        L.pmgo_reset_with.argtypes = [C.c_void_p, dp, dp]
        L.pmgo_step.argtypes = [C.c_void_p, dp, dp, dp, ip, ip]
        L.pmgo_compute_reward.argtypes = [dp, dp, C.c_int64, C.c_int, C.c_double, C.c_int, dp,
                                          C.POINTER(C.c_uint8)]
        L.pmgo_state_size.argtypes = [C.c_void_p]
        L.pmgo_state_size.restype = C.c_int
        L.pmgo_get_state.argtypes = [C.c_void_p, dp]
        L.pmgo_set_state.argtypes = [C.c_void_p, dp]
        L.pmgo_poke_state.argtypes = [C.c_void_p, dp]
        L.pmgo_fk_tip.argtypes = [dp, dp, dp]
        L.pmgo_ik.argtypes = [dp, dp, dp, C.c_int, C.c_double, dp]
        L.pmgo_substeps.argtypes = [C.c_void_p, C.c_int]
        L.pmgo_step_simulation.argtypes = [C.c_void_p]
        L.pmgo_mass_matrix_inverse.argtypes = [C.c_void_p, dp]
        L.pmgo_get_contacts.argtypes = [C.c_void_p, dp, C.c_int]
        L.pmgo_get_contacts.restype = C.c_int
        L.pmgo_link_state.argtypes = [C.c_void_p, C.c_int, dp]
        L.pmgo_rng_uniform.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int, dp]
        L.pmgo_rng_shuffle.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.c_int]
        L.pmgo_bench_rollout.argtypes = [C.POINTER(C.c_void_p), C.c_int, dp, C.c_int, C.c_int]
        L.pmgo_bench_rollout.restype = C.c_double
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def gym_seed_key(seed):
    """gym 0.17.3 utils/seeding.py: np_random(seed) -> RandomState.seed(int list).

    hash_seed = sha512(str(seed))[:8] read as little-endian 32-bit words (zero-padded to 12
    bytes), reassembled into a big int and split again into 32-bit words, dropping leading
    zeros (base_env.py:120-122 calls seeding.np_random).
    """
    seed = int(seed) % 2 ** 64
    h = hashlib.sha512(str(seed).encode("utf8")).digest()[:8]
    h += b"\0" * (4 - len(h) % 4)
    words = struct.unpack("%dI" % (len(h) // 4), h)
    big = sum(2 ** (32 * i) * v for i, v in enumerate(words))
    if big == 0:
        return [0]
    out = []
    while big > 0:
        big, mod = divmod(big, 2 ** 32)
        out.append(mod)
    return out


class OracleEnv:
    """Single-environment CPU oracle with the reference's reset/step semantics."""

    def __init__(self, task="reach", num_block=4, binary_reward=True, distance_threshold=0.05,
                 max_episode_steps=50, seed=0, grip_informed_goal=False, joint_control=False,
                 task_decomposition=False, use_curriculum=False, num_goals_to_generate=10 ** 6):
        self.L = lib()
        self.task = task
        self.h = self.L.pmgo_create_ex3(TASKS[task], num_block, int(binary_reward), distance_threshold,
                                        max_episode_steps, int(grip_informed_goal), int(joint_control),
                                        int(task_decomposition), int(use_curriculum), int(num_goals_to_generate))
        dims = (C.c_int * 4)()
        self.adim = self.L.pmgo_dims(self.h, dims)
        self.dims = list(dims)
        self.nb = 0 if task == "reach" else (num_block if task in ("block_stack", "block_rearrange") else 1)
        self.seed(seed)

    def __del__(self):
        try:
            self.L.pmgo_destroy(self.h)
        except Exception:
            pass

    def seed(self, seed):
        key = gym_seed_key(seed)
        arr = (C.c_uint32 * len(key))(*key)
        self.L.pmgo_seed_array(self.h, arr, len(key))

    def _split(self, flat):
        o, p, a, d = self.dims
        return {"observation": flat[:o].copy(), "policy_state": flat[o:o + p].copy(),
                "achieved_goal": flat[o + p:o + p + a].copy(), "desired_goal": flat[o + p + a:o + p + a + d].copy()}

    def set_sub_goal(self, ind):
        self.L.pmgo_set_sub_goal(self.h, int(ind))

    def set_curriculum_update(self, on):
        self.L.pmgo_set_curriculum_update(self.h, int(on))

    def curriculum(self):
        p = np.zeros(self.nb)
        level = self.L.pmgo_get_curriculum(self.h, _dp(p))
        return p, level

    def observe(self):
        out = np.zeros(sum(self.dims))
        self.L.pmgo_observe(self.h, _dp(out))
        return self._split(out)

    def reset(self):
        out = np.zeros(sum(self.dims))
        self.L.pmgo_reset(self.h, _dp(out))
        return self._split(out)

    def reset_with(self, spawn):
        spawn = np.ascontiguousarray(spawn, dtype=np.float64)
        out = np.zeros(sum(self.dims))
        self.L.pmgo_reset_with(self.h, _dp(spawn), _dp(out))
        return self._split(out)

    def step(self, action):
        a = np.ascontiguousarray(action, dtype=np.float64)
        assert a.shape == (self.adim,)
        out = np.zeros(sum(self.dims))
        r = C.c_double()
        done, ok = C.c_int(), C.c_int()
        self.L.pmgo_step(self.h, _dp(a), _dp(out), C.byref(r), C.byref(done), C.byref(ok))
        return self._split(out), r.value, bool(done.value), {"goal_achieved": bool(ok.value)}

    def get_state(self):
        s = np.zeros(self.L.pmgo_state_size(self.h))
        self.L.pmgo_get_state(self.h, _dp(s))
        return s

    def set_state(self, s):
        s = np.ascontiguousarray(s, dtype=np.float64)
        assert s.shape == (self.L.pmgo_state_size(self.h),)
        self.L.pmgo_set_state(self.h, _dp(s))

    def poke_state(self, s):
        s = np.ascontiguousarray(s, dtype=np.float64)
        assert s.shape == (self.L.pmgo_state_size(self.h),)
        self.L.pmgo_poke_state(self.h, _dp(s))

    def substeps(self, n):
        self.L.pmgo_substeps(self.h, n)

    def step_simulation(self):
        self.L.pmgo_step_simulation(self.h)

    def link_state(self, which):
        o = np.zeros(13)
        self.L.pmgo_link_state(self.h, which, _dp(o))
        return o

    def contacts(self, maxn=128):
        o = np.zeros((maxn, 11))
        n = self.L.pmgo_get_contacts(self.h, _dp(o), maxn)
        return o[:n]

    def minv(self):
        o = np.zeros((9, 9))
        self.L.pmgo_mass_matrix_inverse(self.h, _dp(o))
        return o

    def rng_uniform(self, lo, hi, n):
        o = np.zeros(n)
        self.L.pmgo_rng_uniform(self.h, lo, hi, n, _dp(o))
        return o

    def rng_shuffle(self, n):
        a = np.arange(n, dtype=np.int64)
        self.L.pmgo_rng_shuffle(self.h, a.ctypes.data_as(C.POINTER(C.c_int64)), n)
        return a


def fk_tip(q):
    q = np.ascontiguousarray(q, dtype=np.float64)
    pos, quat = np.zeros(3), np.zeros(4)
    lib().pmgo_fk_tip(_dp(q), _dp(pos), _dp(quat))
    return pos, quat


def ik(q_seed, pos, quat=(0, -1, 0, 0), max_iter=40, thr=1e-5):
    q_seed = np.ascontiguousarray(q_seed, dtype=np.float64)
    pos = np.ascontiguousarray(pos, dtype=np.float64)
    quat = np.ascontiguousarray(quat, dtype=np.float64)
    out = np.zeros(9)
    lib().pmgo_ik(_dp(q_seed), _dp(pos), _dp(quat), max_iter, thr, _dp(out))
    return out


def compute_reward(ag, dg, thr=0.05, binary=True):
    ag = np.ascontiguousarray(ag, dtype=np.float64)
    dg = np.ascontiguousarray(dg, dtype=np.float64)
    g = ag.shape[-1]
    n = ag.size // g
    r = np.zeros(n)
    ok = np.zeros(n, dtype=np.uint8)
    lib().pmgo_compute_reward(_dp(ag), _dp(dg), n, g, thr, int(binary), _dp(r), ok.ctypes.data_as(C.POINTER(C.c_uint8)))
    return r.reshape(ag.shape[:-1]), ok.astype(bool).reshape(ag.shape[:-1])


def bench_rollout(envs, actions, n_threads):
    """actions: [T, n_env, adim] float64. Returns wall seconds."""
    actions = np.ascontiguousarray(actions, dtype=np.float64)
    T, n = actions.shape[0], actions.shape[1]
    hs = (C.c_void_p * n)(*[e.h for e in envs])
    return lib().pmgo_bench_rollout(hs, n, _dp(actions), T, n_threads)
