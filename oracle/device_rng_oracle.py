"""TEST INFRASTRUCTURE -- numpy restatement of the device-side reset sampler (csrc/pmg_spawn.cuh).

The reference samples resets from gym's np_random (MT19937; reproduced bit-exactly by pmg_reset on the host
and by oracle/pmg_oracle.c).  The device path (pmg_reset_device, auto-reset) applies the same sampling rules
(kuka_single_step_base_env.py:104-148, kuka_multi_step_base_env.py:223-240, kuka_multi_step_envs.py:34-87,
174-189) to a counter-based Philox4x32-10 stream (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3",
SC'11; known-answer vectors of the Random123 distribution in tests/test_device_rng.py) in float32.  This file
restates that, one rounding per operation, so that the spawn rows can be compared bit for bit.
Only tests/ import it.
"""
import numpy as np

M0, M1 = 0xD2511F53, 0xCD9E8D57
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF
MAX_TRIES = 64
f32 = np.float32


def philox4x32_10(ctr, key):
    c0, c1, c2, c3 = [int(c) & MASK for c in ctr]
    k0, k1 = [int(k) & MASK for k in key]
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & MASK, p1 & MASK, ((p0 >> 32) ^ c3 ^ k1) & MASK, p0 & MASK
        k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
    return [c0, c1, c2, c3]


class Stream:
    """counter = (block number, episode, env lo, env hi), key = (seed lo, seed hi)"""

    def __init__(self, seed, env, episode):
        self.key = [seed & MASK, (seed >> 32) & MASK]
        self.ctr = [0, episode & MASK, env & MASK, (env >> 32) & MASK]
        self.buf = []

    def u32(self):
        if not self.buf:
            self.buf = philox4x32_10(self.ctr, self.key)
            self.ctr[0] = (self.ctr[0] + 1) & MASK
        return self.buf.pop(0)

    def uniform(self, lo, hi):
        u = f32(self.u32() >> 8) * f32(1.0 / 16777216.0)
        return f32(lo) + (f32(hi) - f32(lo)) * u


def _d2(ax, ay, bx, by):
    dx, dy = f32(ax) - f32(bx), f32(ay) - f32(by)
    return dx * dx + dy * dy


def bounds(task):
    """kuka.py:35-51 with obj_range = target_range = 0.15 (as pmg_create computes them, in double, then float32)."""
    tz = 0.175 + 0.001 if task in ("push", "block_rearrange", "slide") else 0.25
    tip = [-0.52, 0.0, tz]
    obj_range, target_range = (0.1, 0.2) if task == "slide" else (0.15, 0.15)   # kuka_single_step_envs.py:49-59
    obj_lo = [tip[k] - obj_range for k in range(3)]
    obj_hi = [tip[k] + obj_range for k in range(3)]
    tgt_lo = [tip[k] - target_range for k in range(3)]
    tgt_hi = [tip[k] + target_range for k in range(3)]
    obj_lo[0] += 0.03; obj_hi[0] -= 0.03
    tgt_lo[0] += 0.03; tgt_lo[2] = 0.175; tgt_hi[0] -= 0.03
    if task == "slide":
        tgt_lo[0] -= 0.4; tgt_hi[0] -= 0.4                                       # kuka_single_step_base_env.py:66-69
    c = lambda v: [f32(x) for x in v]
    return dict(tip=c(tip), obj_lo=c(obj_lo[:2]), obj_hi=c(obj_hi[:2]), tgt_lo=c(tgt_lo), tgt_hi=c(tgt_hi))


TASK_IDS = {"reach": 0, "push": 1, "pick_and_place": 2, "block_stack": 3, "block_rearrange": 4, "slide": 5}


def sample_row(task, num_block, grip, seed, env, episode):
    """The spawn row [block xy (2 nb) | goal (G)] of one reset, float32."""
    b = bounds(task)
    r = Stream(seed, env, episode)
    t = TASK_IDS[task]
    multi = t in (3, 4)
    nb = 0 if t == 0 else (num_block if multi else 1)
    G = 3 * nb + (4 if grip else 0) if multi else 3
    out = np.zeros(2 * nb + G, dtype=np.float32)
    R01, R006, R008, Z0 = f32(0.01), f32(0.0036), f32(0.0064), (f32(0.17) if t == 5 else f32(0.175))
    if multi:
        for k in range(nb):
            for _ in range(MAX_TRIES):
                x, y = r.uniform(b["obj_lo"][0], b["obj_hi"][0]), r.uniform(b["obj_lo"][1], b["obj_hi"][1])
                ok = _d2(x, y, b["tip"][0], b["tip"][1]) > R006
                for j in range(k):
                    ok = ok and _d2(x, y, out[2 * j], out[2 * j + 1]) > R006
                if ok:
                    break
            out[2 * k], out[2 * k + 1] = x, y
        goal = out[2 * nb:]
        if t == 4:
            for k in range(nb):
                for _ in range(MAX_TRIES):
                    x, y = r.uniform(b["tgt_lo"][0], b["tgt_hi"][0]), r.uniform(b["tgt_lo"][1], b["tgt_hi"][1])
                    ok = True
                    for j in range(k):
                        ok = ok and _d2(x, y, goal[3 * j], goal[3 * j + 1]) > R006
                    for j in range(nb):
                        ok = ok and _d2(x, y, out[2 * j], out[2 * j + 1]) > R006
                    if ok:
                        break
                goal[3 * k:3 * k + 3] = (x, y, Z0)
            return out
        order = list(range(nb))
        for k in range(nb - 1, 0, -1):
            j = (r.u32() * (k + 1)) >> 32
            order[k], order[j] = order[j], order[k]
        for _ in range(MAX_TRIES):
            bx, by = r.uniform(b["tgt_lo"][0], b["tgt_hi"][0]), r.uniform(b["tgt_lo"][1], b["tgt_hi"][1])
            if all(_d2(bx, by, out[2 * j], out[2 * j + 1]) > R008 for j in range(nb)):
                break
        for k in range(nb):
            goal[3 * order[k]:3 * order[k] + 3] = (bx, by, Z0 + f32(0.03) * f32(k))
        if grip:
            goal[3 * nb:3 * nb + 4] = (bx, by, Z0 + f32(0.03) * f32(nb - 1), f32(0.03))
        return out
    cx, cy, cz = b["tip"]
    if nb:
        for _ in range(MAX_TRIES):
            x, y = r.uniform(b["obj_lo"][0], b["obj_hi"][0]), r.uniform(b["obj_lo"][1], b["obj_hi"][1])
            if not (_d2(x, y, b["tip"][0], b["tip"][1]) < R01):
                break
        out[0], out[1] = x, y
        cx, cy, cz = x, y, Z0
    for _ in range(MAX_TRIES):
        g = [r.uniform(b["tgt_lo"][k], b["tgt_hi"][k]) for k in range(3)]
        dz = f32(g[2]) - f32(cz)
        if _d2(g[0], g[1], cx, cy) + dz * dz > R01:
            break
    if t in (1, 5):
        g[2] = Z0
    elif t == 2:
        if r.uniform(0.0, 1.0) >= f32(0.5):
            g[2] = Z0
    out[2 * nb:2 * nb + 3] = g
    return out
