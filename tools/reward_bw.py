"""_compute_reward over [n, g] rows (pmg_compute_reward): correctness against torch on ragged / unaligned shapes and
achieved HBM bandwidth (algorithmic bytes 8g + 5 per row, inputs larger than L2).  PMG_REWARD_SIMPLE=1 selects the
one-thread-per-row kernel for A/B."""
import ctypes as C
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pybullet_multigoal_gym_b200 import _lib

L = _lib.load()
dev = torch.device("cuda:0")
ptr = lambda t: C.c_void_p(t.data_ptr())
thr = 0.05


def call(ag, dg, n, g, binary, r, ok):
    _lib.check(L.pmg_compute_reward(ptr(ag), ptr(dg), C.c_int64(n), C.c_int32(g), C.c_float(thr), C.c_int32(binary), ptr(r), ptr(ok), C.c_void_p(0)))


torch.manual_seed(0)
bad = 0
for n, g, off in [(1, 3, 0), (255, 3, 0), (1025, 3, 0), (2051, 1, 0), (700, 6, 0), (513, 12, 0), (300, 15, 0), (257, 16, 0), (1000, 7, 0),
                  (33, 32, 0), (40, 33, 0), (1000003, 3, 0), (100001, 12, 0), (5000, 3, 1), (5000, 12, 3)]:
    flat_a = torch.rand(n * g + off, device=dev) * 0.1
    flat_b = flat_a + (torch.rand(n * g + off, device=dev) - 0.5) * 0.08
    ag, dg = flat_a[off:], flat_b[off:]          # off != 0: not 16-byte aligned -> the row-per-thread kernel
    for binary in (1, 0):
        r = torch.empty(n, device=dev); ok = torch.empty(n, dtype=torch.uint8, device=dev)
        call(ag, dg, n, g, binary, r, ok)
        d = (ag.view(n, g).double() - dg.view(n, g).double()).norm(dim=1)
        safe = (d - thr).abs() > 1e-6
        want_ok = d <= thr
        e_ok = int((ok.bool() != want_ok)[safe].sum())
        want_r = -(~want_ok).double() if binary else -d
        e_r = float(((r.double() - want_r).abs() * (safe if binary else torch.ones_like(safe))).max())
        if e_ok or e_r > 1e-6:
            bad += 1
            print("MISMATCH n=%d g=%d off=%d binary=%d: flags %d, reward %.3g" % (n, g, off, binary, e_ok, e_r))
print("correctness: %s" % ("all shapes agree with torch (flags away from the threshold, rewards to 1e-6)" if not bad else "%d mismatches" % bad))

peak = 6542.4
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
for g, n in [(3, 1 << 26), (12, 1 << 25), (16, 1 << 24)]:
    ag = torch.rand(n * g, device=dev); dg = torch.rand(n * g, device=dev)
    r = torch.empty(n, device=dev); ok = torch.empty(n, dtype=torch.uint8, device=dev)
    for _ in range(3):
        call(ag, dg, n, g, 1, r, ok)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        call(ag, dg, n, g, 1, r, ok)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    gbs = n * (8 * g + 5) / (ms / 1e3) / 1e9
    print("g=%2d n=%9d (%.2f GB per launch, > L2): %.3f ms  %.0f GB/s = %.1f %% of the measured %.0f GB/s  (%.2f G rows/s)" % (
        g, n, n * (8 * g + 5) / 1e9, ms, gbs, 100 * gbs / peak, peak, n / ms / 1e6))
    del ag, dg, r, ok
