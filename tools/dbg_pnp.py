import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pybullet_multigoal_gym_b200 as pmg
from oracle import pmg_oracle as O
np.set_printoptions(precision=5, suppress=True, linewidth=220)
N = 8
env = pmg.make_env(task="pick_and_place", batch=N, max_episode_steps=80)
gobs = env.reset()
refs = []
for seed in range(N):
    e = O.OracleEnv('pick_and_place', max_episode_steps=80, seed=seed); e.reset(); refs.append(e)
obs = [e.reset() for e in refs]
print("reset diff", max(np.abs(gobs['observation'][i].cpu().numpy() - obs[i]['observation']).max() for i in range(N)))
phase = [0]*N; timer=[0]*N
first_div = [None]*N
for t in range(70):
    A = np.zeros((N,4), dtype=np.float32)
    for i in range(N):
        tip=obs[i]['observation'][:3]; blk=obs[i]['achieved_goal']; goal=obs[i]['desired_goal']
        tgt = blk+[0,0,0.06] if phase[i]==0 else (goal if phase[i]==3 else blk)
        a=np.zeros(4); a[:3]=np.clip((tgt-tip)/0.01,-1,1); a[3]= 1.0 if phase[i]>=2 else -1.0
        if phase[i]==2: a[:3]=0; timer[i]+=1
        err=np.linalg.norm(tgt-tip)
        if phase[i]==0 and err<0.008: phase[i]=1
        elif phase[i]==1 and err<0.004: phase[i]=2
        elif phase[i]==2 and timer[i]>=4: phase[i]=3
        A[i]=a
    g = env.step(torch.from_numpy(A).cuda())[0]
    for i in range(N):
        obs[i] = refs[i].step(A[i].astype(np.float64))[0]
        d = np.abs(g['observation'][i].cpu().numpy() - obs[i]['observation'])
        if first_div[i] is None and d[:10].max() > 1e-3:
            first_div[i] = t
            print("env", i, "diverges at step", t, "phase", phase[i], "\n gpu", g['observation'][i].cpu().numpy()[:14], "\n ora", obs[i]['observation'][:14])
print("first divergence steps", first_div, "phases", phase)
print("overflow", env.overflow_count)
st = env.get_state()
print("gpu final block z", st[:,48], "oracle", [o['achieved_goal'][2] for o in obs])
