import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, contextlib, io
import pybullet_multigoal_gym_b200 as pmg
from oracle import pmg_oracle as O
KEYS = ("observation", "policy_state", "achieved_goal", "desired_goal")
kw = dict(task="block_stack", num_block=3, grip_informed_goal=True)
B = 6
with contextlib.redirect_stdout(io.StringIO()):
    env = pmg.make_env(batch=B, **kw)
env.reset(); spawn = env.last_spawn()
refs = []
for i in range(B):
    o = O.OracleEnv(seed=i, **kw); o.reset_with(spawn[i].astype(np.float64)); refs.append(o)
rng = np.random.RandomState(11)
for t in range(24):
    st = np.stack([o.get_state() for o in refs]).astype(np.float32)
    a = rng.uniform(-1, 1, size=(B, 4)).astype(np.float32)
    for i in range(B):
        refs[i].set_state(st[i].astype(np.float64)); rng.randn(9)
        tip = refs[i].link_state(0)[:3]
        a[i, :3] = np.clip((st[i, 46:49] + np.array([0.0, 0.0, 0.0 if t > 8 else 0.06]) - tip) / 0.01, -1, 1)
        a[i, 3] = -1.0 if t < 14 else 1.0
    env.set_state(st)
    obs, r, done, info = env.step(torch.from_numpy(a).cuda())
    got = np.concatenate([obs[k].cpu().numpy() for k in KEYS], axis=1)
    gst = env.get_state()
    for i in range(B):
        ro = refs[i].step(a[i].astype(np.float64))[0]
        want = np.concatenate([ro[k] for k in KEYS])
        d = np.abs(got[i] - want)
        cols = np.r_[0:3, want.size - 2 * env.goal_dim:want.size]
        mask = np.zeros_like(d); mask[cols] = 1; d = d * mask
        ost = refs[i].get_state()
        if d.max() > 1e-4:
            j = int(np.argmax(d))
            print("t", t, "env", i, "col", j, "err %.3g" % d.max(), "got", got[i, j], "want", want[j], "| q78 gpu", gst[i, 7:9], "oracle", ost[7:9], "contacts", len(refs[i].contacts()))
