"""Debug aid: the scripted PickAndPlace teacher-forced steps on the GPU, the emulator (same source on the CPU) and
the oracle side by side; prints the velocity entries where they disagree."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import pmg_oracle as O
from tests.emu import build_emu
import pybullet_multigoal_gym_b200 as pmg
O.build()
emu = C.CDLL(build_emu.build())
FP = C.POINTER(C.c_float); U8 = C.POINTER(C.c_uint8)
_f = lambda a: a.ctypes.data_as(FP)
KEYS = ("observation", "policy_state", "achieved_goal", "desired_goal")
SCR = [(10, (0.0, 0.0, 0.07), -1.0), (10, (0.0, 0.0, 0.0), -1.0), (5, (0.0, 0.0, 0.0), 1.0), (12, (0.0, 0.0, 0.10), 1.0)]
phases = []
for n, rel, grip in SCR: phases += [(np.array(rel), grip)] * n
B = 8
env = pmg.make_env(task="pick_and_place", batch=B, binary_reward=False)
env.reset()
spawn = env.last_spawn()
refs = []
for i in range(B):
    o = O.OracleEnv("pick_and_place", seed=i, binary_reward=False); o.reset_with(spawn[i].astype(np.float64)); refs.append(o)
for t, (rel, grip) in enumerate(phases):
    st = np.stack([o.get_state() for o in refs]).astype(np.float32)
    a = np.zeros((B, 4), np.float32)
    for i in range(B):
        refs[i].set_state(st[i].astype(np.float64))
        tip = refs[i].link_state(0)[:3]
        a[i, :3] = np.clip((st[i, 46:49] + rel - tip) / 0.01, -1, 1); a[i, 3] = grip
    env.set_state(st)
    obs, r, d, info = env.step(torch.from_numpy(a).cuda())
    got = np.concatenate([obs[k].cpu().numpy() for k in KEYS], axis=1)
    st_gpu = env.get_state()
    for i in range(B):
        want = np.concatenate([refs[i].step(a[i].astype(np.float64))[0][k] for k in KEYS])
        man = np.zeros(6 * 41, np.float32); eo, rew = np.zeros(33, np.float32), np.zeros(1, np.float32); dn, su = np.zeros(1, np.uint8), np.zeros(1, np.uint8)
        s2 = st[i].copy()
        emu.pmg_emu_block_step(2, _f(s2), _f(man), _f(a[i].copy()), C.c_float(0.05), 0, 50, _f(eo), _f(rew), dn.ctypes.data_as(U8), su.ctypes.data_as(U8))
        if np.abs(got[i] - want)[10:20].max() > 5e-3 or np.abs(eo - got[i])[10:20].max() > 5e-3:
            np.set_printoptions(precision=4, suppress=True, linewidth=200)
            print("t=%d env=%d\n  gpu    %s\n  emu    %s\n  oracle %s" % (t, i, got[i][10:20], eo[10:20], want[10:20]))
            print("  state after: gpu q %s qd %s\n               emu q %s qd %s\n            oracle q %s qd %s" % (st_gpu[i][7:9], st_gpu[i][16:18], s2[7:9], s2[16:18], refs[i].get_state()[7:9], refs[i].get_state()[16:18]))
            print("  block w: gpu %s emu %s oracle %s" % (st_gpu[i][56:59], s2[56:59], refs[i].get_state()[56:59]))
print("overflow", env.overflow_count)
