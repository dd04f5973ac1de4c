"""Development aid: where the host-buffer path (env.step(numpy) -> pmg_step_host) spends its time."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pybullet_multigoal_gym_b200 as pmg
from pybullet_multigoal_gym_b200 import _lib
from pybullet_multigoal_gym_b200.envs import _ptr
B = 8192
env = pmg.make_env(task="reach", batch=B, check_actions=False)
L = _lib.load()
rng = np.random.RandomState(0)
acts = rng.uniform(-1, 1, size=(50, B, 3)).astype(np.float32)
dev = torch.rand((50, B, 3), device="cuda") * 2 - 1
out = torch.empty((B, env.row_width), device="cuda"); r = torch.empty((B,), device="cuda")
d = torch.empty((B,), dtype=torch.uint8, device="cuda"); s = torch.empty((B,), dtype=torch.uint8, device="cuda")
def run(name, fn):
    env.reset(); fn(0); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(50): fn(k)
    torch.cuda.synchronize()
    print("%-46s %.3f ms/step" % (name, (time.perf_counter() - t0) / 50 * 1e3))
run("device path, step_packed + one sync at the end", lambda k: env.step_packed(dev[k], out, r, d, s))
def dev_sync(k):
    env.step_packed(dev[k], out, r, d, s); torch.cuda.synchronize()
run("device path, sync every step", dev_sync)
st = env._stream()
def raw_host(k):
    env._h_action.numpy()[...] = acts[k]
    L.pmg_step_host(env._h, _ptr(env._h_action), _ptr(env._h_obs), _ptr(env._h_reward), _ptr(env._h_done), _ptr(env._h_success), st)
run("pmg_step_host through ctypes (no dict assembly)", raw_host)
run("env.step(numpy)", lambda k: env.step(acts[k]))
