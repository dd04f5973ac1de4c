"""Teacher-forced one-step parity of the CUDA step kernels against the CPU oracle (shared by the GPU tests).

Contact-rich rollouts are chaotic in open loop (SURVEY.md 7 hard part 3), so every env.step starts from the
oracle's own fp32-rounded state.  Stiff contact events amplify even a 1e-7 perturbation inside the
double-precision oracle itself into millimetres (a jaw landing on a block edge: the oracle is tri-modal there),
so N_TWINS perturbed copies of the oracle measure that sensitivity per step -- one twin misses such a bifurcation
about one time in five, which is what the former 1 cm escape hatch of these tests papered over.

Every entry of the packed row [observation | policy_state | achieved_goal | desired_goal] is compared, in two
classes: POSITION entries (tip, block positions / quaternions, jaw opening, goals, joint poses) and VELOCITY
entries (tip / finger velocity, relative linear / angular block velocities).  The contact ERP term (0.9 / 2 ms)
amplifies fp32 rounding of a contact depth 450-fold into velocity, so velocity entries of bodies in sliding
contact carry up to a few 1e-3 m/s | rad/s of noise where positions stay within 1e-4; the fraction of env-steps
within 1e-4 is printed for both classes.
"""
import numpy as np
import torch

KEYS = ("observation", "policy_state", "achieved_goal", "desired_goal")
TOL = 1e-4
N_TWINS = 4


def velocity_mask(task, num_block, row_width, joint_control=False):
    """Boolean mask over the packed row: True for velocity entries (kuka_single_step_base_env.py:208-209,
    kuka_multi_step_base_env.py:276-283; joint control prepends the 7 joint poses, :214-216)."""
    vel = np.zeros(row_width, dtype=bool)
    jo = 7 if joint_control else 0
    if task in ("push", "pick_and_place"):
        vel[jo + 10:jo + 20] = True          # tip vel 3, finger vel 1, rel lin vel 3, rel ang vel 3
    elif task in ("block_stack", "block_rearrange"):
        vel[jo + 4:jo + 8] = True            # tip vel 3, finger vel 1
        for n in range(num_block):
            vel[jo + 8 + 16 * n + 10:jo + 8 + 16 * n + 16] = True
    return vel


class Stats:
    def __init__(self, name):
        self.name, self.pos, self.vel, self.loose = name, [], [], 0

    def report(self):
        pos, vel = np.array(self.pos), np.array(self.vel)
        msg = ("%s teacher-forced: %d well-conditioned env-steps (%d ill-conditioned by the oracle's own %d-twin sensitivity); "
               "position entries: %.1f%% of env-steps within 1e-4, median %.2g, worst %.3g" %
               (self.name, pos.size, self.loose, N_TWINS, 100 * float(np.mean(pos < TOL)), np.median(pos), pos.max()))
        if vel.size and vel.max() > 0:
            msg += ("; velocity entries: %.1f%% within 1e-4, %.1f%% within 1e-3, worst %.3g" %
                    (100 * float(np.mean(vel < TOL)), 100 * float(np.mean(vel < 1e-3)), vel.max()))
        print(msg)
        return pos, vel


def run(env, oracle, refs, twins, nsteps, action_fn, vel, rng, name, envs=None, perturb_block=False):
    """refs[j] / twins[j][k]: oracle envs for batch index envs[j] (default: all).  action_fn(t, j, state_row, tip)
    -> action row.  Returns Stats; asserts the hard bounds: well-conditioned position entries < 5e-4, velocity
    entries < 2e-2, ill-conditioned steps within 50x the oracle's own sensitivity."""
    B = env.batch
    idx = np.arange(B) if envs is None else np.asarray(envs)
    stats = Stats(name)
    A = env.action_dim
    for t in range(nsteps):
        st = np.stack([o.get_state() for o in refs]).astype(np.float32)
        a = rng.uniform(-1, 1, size=(B, A)).astype(np.float32)
        for j, o in enumerate(refs):
            o.set_state(st[j].astype(np.float64))
            for tw in twins[j]:
                pert = st[j].astype(np.float64)
                pert[:9] += 1e-7 * rng.randn(9)
                if perturb_block:
                    pert[46:49] += 1e-7 * rng.randn(3)
                tw.set_state(pert)
            a[idx[j]] = action_fn(t, j, st[j], o.link_state(0)[:3], a[idx[j]])
        if envs is None:
            env.set_state(st)
        else:
            full = env.get_state()
            full[idx] = st
            env.set_state(full)
        obs, r, done, info = env.step(torch.from_numpy(a).cuda())
        got = np.concatenate([obs[k].detach().cpu().numpy() for k in KEYS], axis=1)
        for j, o in enumerate(refs):
            aj = a[idx[j]].astype(np.float64)
            ro = o.step(aj)[0]
            want = np.concatenate([ro[k] for k in KEYS])
            sens_p = sens_v = 0.0
            for tw in twins[j]:
                rt = tw.step(aj)[0]
                d = np.abs(np.concatenate([rt[k] for k in KEYS]) - want)
                sens_p = max(sens_p, float(d[~vel].max()))
                sens_v = max(sens_v, float(d[vel].max()) if vel.any() else 0.0)
            d = np.abs(got[idx[j]] - want)
            err_p, err_v = float(d[~vel].max()), (float(d[vel].max()) if vel.any() else 0.0)
            if sens_p < 2e-6:
                stats.pos.append(err_p)
                stats.vel.append(err_v)
                assert err_p < 5 * TOL, (name, t, int(idx[j]), err_p, sens_p)
                assert err_v < max(2e-2, 50 * sens_v), (name, t, int(idx[j]), err_v, sens_v)
            else:
                stats.loose += 1
                assert err_p < max(50 * sens_p, 10 * TOL), (name, t, int(idx[j]), err_p, sens_p)
    return stats
