"""Batched Kuka multigoal environments behind the reference's env API.

Mirrors (paths relative to /root/reference/pybullet_multigoal_gym/):
  envs/base_envs/base_env.py:120-138          seed / reset / step
  envs/base_envs/kuka_single_step_base_env.py:193-244   observation dict, _compute_reward
  envs/base_envs/kuka_multi_step_base_env.py:255-345    multi-block observation dict, reward
  envs/task_envs/kuka_single_step_envs.py:4-59, kuka_multi_step_envs.py:6-32,151-189   task presets
  robots/kuka.py:104-118,204-206                joint-space control variant (joint_control=True)
  envs/base_envs/kuka_multi_step_base_env.py:300-304   grip-informed goals (grip_informed_goal=True)
  gym 0.17.3 wrappers/time_limit.py           done = elapsed >= max_episode_steps

Every array of the reference gains a leading [batch] axis (unless batch=None, which gives the
reference's unbatched numpy shapes for a single environment).  All physics runs in the CUDA
extension (libpmg.so); there is no CPU path.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib, seeding, spaces

TASK_IDS = {"reach": 0, "push": 1, "pick_and_place": 2, "block_stack": 3, "block_rearrange": 4, "slide": 5}


class StepDemonstrator:
    """Cycles through the sub-goal indices of a demonstration, same interface as the reference's
    utils/demonstrator.py:1-35 (get_next_goal / manual_reset / reset_with_the_last_sub_goal_index)."""

    def __init__(self, demonstrations, stick_with_final_goal=True):
        self.demonstrations, self.demon_num = demonstrations, len(demonstrations)
        self.demon_ind, self.current_goal, self.current_final_goal = 0, -1, 0
        self.stick_with_final_goal, self.final = stick_with_final_goal, False

    def get_next_goal(self):
        demo = self.demonstrations[self.demon_ind]
        if self.stick_with_final_goal and self.current_goal != -1:
            self.final = False
            if demo[self.current_goal] == demo[-1]:
                self.final = True
                return demo[self.current_goal]
        self.current_goal = (self.current_goal + 1) % len(demo)
        return demo[self.current_goal]

    def manual_reset(self, demon_ind=None):
        self.demon_ind = 0 if demon_ind is None else demon_ind
        self.current_goal, self.final = -1, False
        self.current_final_goal = self.demonstrations[self.demon_ind][-1]

    def reset_with_the_last_sub_goal_index(self, ind):
        self.current_goal = -1
        for i, demo in enumerate(self.demonstrations):
            if demo[-1] == ind:
                self.demon_ind = i
                break
        self.current_final_goal = self.demonstrations[self.demon_ind][-1]
        self.final = False


class ActionError(AssertionError, ValueError):
    """Raised where the reference hits `assert self.action_space.contains(a)` (kuka.py:168)."""


def _ptr(t):
    return C.c_void_p(t.data_ptr())


class KukaBulletMGEnv:
    """Batched goal-conditioned Kuka environment (one CUDA thread per environment)."""

    metadata = {"render.modes": []}

    def __init__(self, task, batch=None, device=0, binary_reward=True, distance_threshold=0.05,
                 max_episode_steps=50, num_block=4, seed=0, check_actions=True,
                 grip_informed_goal=False, joint_control=False, task_decomposition=False,
                 use_curriculum=False, num_goals_to_generate=1e6, device_sampling=False, auto_reset=False,
                 env_index_base=0):
        if task not in TASK_IDS:
            raise ValueError("invalid task name: %s, only support: %s" % (task, sorted(TASK_IDS)))
        self._L = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("pybullet_multigoal_gym_b200 needs a CUDA device (no CPU fallback)")
        self.task = task
        self._squeeze = batch is None
        self.batch = self.num_envs = 1 if batch is None else int(batch)
        self.device = torch.device("cuda", int(device) if not isinstance(device, torch.device) else device.index or 0)
        self.binary_reward = bool(binary_reward)
        self.distance_threshold = float(distance_threshold)
        self._max_episode_steps = int(max_episode_steps)
        multi = task in ("block_stack", "block_rearrange")
        self.num_block = int(num_block) if multi else (0 if task == "reach" else 1)
        self.grasping = task in ("pick_and_place", "block_stack")
        self.has_obj = task != "reach"
        self.joint_control = bool(joint_control)
        self.grip_informed_goal = bool(grip_informed_goal)
        if self.grip_informed_goal and task != "block_stack":
            # kuka_multi_step_envs.py:158 asserts it off for rearrange; the single-step envs have no such option
            raise AssertionError("%s does not support gripper informed goal representation." % task)
        self.task_decomposition = bool(task_decomposition) and task == "block_stack"
        if self.task_decomposition:
            # kuka_multi_step_envs.py:13-17, kuka_multi_step_base_env.py:116-119: demonstrations [0], [0, 1], ...
            self.num_steps = self.num_block * (2 if self.grip_informed_goal else 1)
            self.step_demonstrator = StepDemonstrator([list(range(i + 1)) for i in range(self.num_steps)])
        self.curriculum = bool(use_curriculum) and task in ("block_stack", "block_rearrange")
        if self.curriculum:
            # kuka_multi_step_base_env.py:112-140
            assert not self.task_decomposition, 'if using curriculum, task decomposition should be False, vice versa'
            import warnings
            warnings.warn("You will need to call env.activate_curriculum_update() before your training phase, "
                          "and env.deactivate_curriculum_update() before your evaluation phase.")
            self.curriculum_update = False
            self.num_curriculum = self.num_block
            self.base_curriculum_episode_steps = 50
            self.num_goals_per_curriculum = int(num_goals_to_generate) // self.num_curriculum
        self.check_actions = check_actions
        cfg = _lib.PmgConfig(TASK_IDS[task], int(num_block), self.batch, int(self.binary_reward),
                             self.distance_threshold, self._max_episode_steps, self.device.index,
                             int(self.grip_informed_goal), int(self.joint_control), int(self.task_decomposition),
                             int(self.curriculum), int(min(int(num_goals_to_generate), 2 ** 31 - 1)))
        h = C.c_void_p()
        _lib.check(self._L.pmg_create(C.byref(cfg), C.byref(h)))
        self._h = h
        dims = (C.c_int32 * 6)()
        _lib.check(self._L.pmg_dims(self._h, dims))
        self.obs_dim, self.policy_dim, self.goal_dim, _, self.action_dim, self.row_width = list(dims)
        self._slices = {
            "observation": slice(0, self.obs_dim),
            "policy_state": slice(self.obs_dim, self.obs_dim + self.policy_dim),
            "achieved_goal": slice(self.obs_dim + self.policy_dim, self.obs_dim + self.policy_dim + self.goal_dim),
            "desired_goal": slice(self.obs_dim + self.policy_dim + self.goal_dim, self.row_width),
        }
        B = self.batch
        # pinned host staging for the host-buffer path (pmg_step_host)
        self._h_action = torch.empty((B, self.action_dim), dtype=torch.float32).pin_memory()
        self._h_obs = torch.empty((B, self.row_width), dtype=torch.float32).pin_memory()
        self._h_reward = torch.empty((B,), dtype=torch.float32).pin_memory()
        self._h_done = torch.empty((B,), dtype=torch.uint8).pin_memory()
        self._h_success = torch.empty((B,), dtype=torch.uint8).pin_memory()
        self._h_blocks = torch.empty((B * self.row_width,), dtype=torch.float32).pin_memory()
        # numpy views and raw pointers of the pinned staging buffers, built once (env.step is called at kHz rates)
        self._np_action, self._np_reward = self._h_action.numpy(), self._h_reward.numpy()
        self._np_done, self._np_success = self._h_done.numpy(), self._h_success.numpy()
        blocks, off = self._h_blocks.numpy(), 0
        self._np_blocks = {}
        for key, dim in (("observation", self.obs_dim), ("policy_state", self.policy_dim),
                         ("achieved_goal", self.goal_dim), ("desired_goal", self.goal_dim)):
            self._np_blocks[key] = blocks[off:off + B * dim].reshape(B, dim)
            off += B * dim
        self._p_action, self._p_blocks, self._p_reward = _ptr(self._h_action), _ptr(self._h_blocks), _ptr(self._h_reward)
        self._p_done, self._p_success = _ptr(self._h_done), _ptr(self._h_success)
        self.action_space = spaces.Box(-np.ones([self.action_dim]), np.ones([self.action_dim]))  # kuka.py:109-118
        self.desired_goal = None
        self.device_sampling = self.auto_reset = False
        self._terminal = None
        self.seed(seed)
        if device_sampling or auto_reset:
            self.enable_device_sampling(seed=seed, env_index_base=env_index_base)
        obs = self.reset()  # the reference ctor resets once (base_env.py:84) and so consumes the RNG
        if auto_reset:
            self.set_auto_reset(True)
        shp = (lambda k: tuple(obs[k].shape))
        self.observation_space = spaces.Dict(dict(  # base_env.py:86-92 (note: 'state', not 'observation')
            state=spaces.Box(-np.inf, np.inf, shape=shp("observation"), dtype="float32"),
            policy_state=spaces.Box(-np.inf, np.inf, shape=shp("policy_state"), dtype="float32"),
            achieved_goal=spaces.Box(-np.inf, np.inf, shape=shp("achieved_goal"), dtype="float32"),
            desired_goal=spaces.Box(-np.inf, np.inf, shape=shp("desired_goal"), dtype="float32"),
        ))

    # ---- gym.Env surface ------------------------------------------------------------------
    def seed(self, seed=None):
        """base_env.py:120-122.  Environment i of the batch is seeded with seed + i, so env 0
        reproduces the reference's stream for `seed` and the others are decorrelated."""
        if seed is None:
            seed = int(np.random.SeedSequence().generate_state(1)[0])
        keys = [seeding.seed_key(int(seed) + i) for i in range(self.batch)]
        width = max(len(k) for k in keys)
        arr = np.zeros((self.batch, width), dtype=np.uint32)
        lens = np.zeros((self.batch,), dtype=np.int32)
        for i, k in enumerate(keys):
            arr[i, :len(k)] = k
            lens[i] = len(k)
        _lib.check(self._L.pmg_seed(self._h, arr.ctypes.data_as(C.c_void_p), lens.ctypes.data_as(C.c_void_p), width))
        return [seed]

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _split(self, packed):
        return {k: packed[:, s] for k, s in self._slices.items()}

    def _to_host_obs(self, packed_np):
        obs = {k: packed_np[:, s].copy() for k, s in self._slices.items()}
        if self._squeeze:
            obs = {k: v[0].astype(np.float64) for k, v in obs.items()}
        return obs

    def reset(self, test=False, mask=None, spawn=None, device_output=None):
        """base_env.py:124-128.  `mask` ([batch] bool) resets a subset; `spawn` ([batch, 2*nb+G])
        overrides the sampled block xy / goal.  Returns torch CUDA tensors when the env is batched
        (numpy for batch=None) unless device_output says otherwise."""
        if self.device_sampling and spawn is None:
            return self._reset_device(mask, device_output)
        with torch.cuda.device(self.device):
            out = torch.empty((self.batch, self.row_width), dtype=torch.float32, device=self.device)
            m = None
            if mask is not None:
                m = np.ascontiguousarray(np.asarray(mask.cpu() if torch.is_tensor(mask) else mask), dtype=np.uint8)
                if m.shape != (self.batch,):
                    raise ValueError("mask must have shape (%d,)" % self.batch)
            sp = None
            if spawn is not None:
                sp = np.ascontiguousarray(np.asarray(spawn.cpu() if torch.is_tensor(spawn) else spawn), dtype=np.float32)
                if sp.shape != (self.batch, self._L.pmg_spawn_width(self._h)):
                    raise ValueError("spawn must have shape (%d, %d)" % (self.batch, self._L.pmg_spawn_width(self._h)))
            _lib.check(self._L.pmg_reset(self._h, m.ctypes.data_as(C.c_void_p) if m is not None else None,
                                         sp.ctypes.data_as(C.c_void_p) if sp is not None else None, _ptr(out), self._stream()))
            obs = self._split(out)
            self.desired_goal = obs["desired_goal"]
            if device_output is None:
                device_output = not self._squeeze
            if device_output:
                return obs
            host = self._to_host_obs(out.cpu().numpy())
            if self._squeeze:
                self.desired_goal = host["desired_goal"]
            return host

    # ---- device-side sampling / auto-reset (no reference counterpart; include/pmg.h) -------------------
    def enable_device_sampling(self, seed=0, env_index_base=0):
        """Resets are sampled inside the reset kernel from Philox streams keyed by (seed, env_index_base + i, episode)
        instead of on the host from the reference's MT19937 streams: same sampling rules, different random numbers,
        no host work and no synchronisation (pmg_set_device_rng)."""
        _lib.check(self._L.pmg_set_device_rng(self._h, int(seed) & (2 ** 64 - 1), int(env_index_base)))
        self.device_sampling = True

    def set_auto_reset(self, on=True, keep_terminal_observation=False):
        """gym VectorEnv-style auto-reset on the device: after every step the environments whose episode ended
        (`done`) reset themselves and their observation rows are those of the new episode; reward / done / info are the
        terminal step's.  With keep_terminal_observation the terminal rows are returned in
        info['terminal_observation'] (valid where done)."""
        if on and not self.device_sampling:
            self.enable_device_sampling()
        self._terminal = None
        if on and keep_terminal_observation:
            self._terminal = torch.zeros((self.batch, self.row_width), dtype=torch.float32, device=self.device)
        _lib.check(self._L.pmg_set_auto_reset(self._h, int(bool(on)), _ptr(self._terminal) if self._terminal is not None else None))
        self.auto_reset = bool(on)

    def _reset_device(self, mask, device_output):
        with torch.cuda.device(self.device):
            out = torch.empty((self.batch, self.row_width), dtype=torch.float32, device=self.device)
            m = None
            if mask is not None:
                m = torch.as_tensor(mask).to(self.device).ne(0).to(torch.uint8).contiguous()
                if tuple(m.shape) != (self.batch,):
                    raise ValueError("mask must have shape (%d,)" % self.batch)
            _lib.check(self._L.pmg_reset_device(self._h, _ptr(m) if m is not None else None, _ptr(out), self._stream()))
            obs = self._split(out)
            self.desired_goal = obs["desired_goal"]
            if device_output is None:
                device_output = not self._squeeze
            if device_output:
                return obs
            host = self._to_host_obs(out.cpu().numpy())
            if self._squeeze:
                self.desired_goal = host["desired_goal"]
            return host

    # ---- curriculum (kuka_multi_step_base_env.py:147-157, 350-379) -------------------------------------
    def _curriculum_toggle(self, on):
        if not self.curriculum:
            import warnings
            warnings.warn("This method should not be called while not using curriculum.")
            return
        self.curriculum_update = on
        _lib.check(self._L.pmg_set_curriculum_update(self._h, int(on)))

    def activate_curriculum_update(self):
        self._curriculum_toggle(True)

    def deactivate_curriculum_update(self):
        self._curriculum_toggle(False)

    def _curriculum_state(self):
        prob = np.zeros((self.batch, self.num_block), dtype=np.float32)
        level = np.zeros((self.batch,), dtype=np.int32)
        _lib.check(self._L.pmg_get_curriculum(self._h, prob.ctypes.data_as(C.c_void_p), level.ctypes.data_as(C.c_void_p)))
        return prob, level

    @property
    def curriculum_prob(self):
        """[batch, num_curriculum]: every environment runs its own schedule, like separate reference envs."""
        p = self._curriculum_state()[0]
        return p[0].astype(np.float64) if self._squeeze else p

    @property
    def last_curriculum_level(self):
        lv = self._curriculum_state()[1]
        return int(lv[0]) if self._squeeze else lv

    @property
    def last_ind_block_to_move(self):
        """block_rearrange with the curriculum (kuka_multi_step_envs.py:201-205): the sorted indices of the blocks the
        last reset gave targets to; a list (single env) or one list per environment."""
        if not (self.curriculum and self.task == "block_rearrange"):
            return None
        masks = self.last_spawn()[:, -1].astype(np.int64)
        moved = [[b for b in range(self.num_block) if (int(m) >> b) & 1] for m in masks]
        return moved[0] if self._squeeze else moved

    @property
    def curriculum_goal_step(self):
        """kuka_multi_step_envs.py:128: the number of episode steps the reference suggests for the drawn level."""
        return self.last_curriculum_level * 25 + self.base_curriculum_episode_steps

    def set_sub_goal(self, sub_goal_ind):
        """kuka_multi_step_base_env.py:159-181.  `sub_goal_ind`: one index for every environment or a [batch]
        array, python list indexing (-1 = the final goal).  Returns the new desired goal like the reference
        (None, with the reference's warning, when the env was made without task_decomposition)."""
        if not self.task_decomposition:
            import warnings
            warnings.warn("The set_sub_goal() method should only be called when using task decomposition,\n"
                          "It does nothing and returns None when self.task_decomposition is False.")
            return None
        ind = np.ascontiguousarray(np.broadcast_to(np.asarray(sub_goal_ind.cpu() if torch.is_tensor(sub_goal_ind) else sub_goal_ind,
                                                              dtype=np.int32), (self.batch,)))
        with torch.cuda.device(self.device):
            _lib.check(self._L.pmg_set_sub_goal(self._h, ind.ctypes.data_as(C.c_void_p), self._stream()))
        obs = self.reset(mask=np.zeros((self.batch,), dtype=np.uint8))  # no env is reset: the observation rows are rebuilt
        return obs["desired_goal"]

    def _check_action(self, a_np):
        if a_np.shape != (self.batch, self.action_dim) or not (np.all(a_np >= -1.0) and np.all(a_np <= 1.0)):
            raise ActionError("action outside the action space Box(-1, 1, (%d,))" % self.action_dim)

    def step(self, action):
        """base_env.py:130-138 + TimeLimit.  CUDA tensor in -> CUDA tensors out, asynchronous on the current stream: the
        action-space check of the reference (kuka.py:168) runs inside the step kernel, which raises a flag in mapped
        host memory; it is looked at here WITHOUT synchronising, so an out-of-range action raises ActionError at the
        first later step / `action_error()` call that finds its kernel finished (the way CUDA reports its own errors).
        numpy / CPU tensor in -> numpy out through the host-buffer C-ABI call, checked before anything is launched."""
        if torch.is_tensor(action) and action.is_cuda:
            return self._step_device(action)
        a = np.asarray(action.numpy() if torch.is_tensor(action) else action, dtype=np.float32)
        if self._squeeze:
            if a.shape != (self.action_dim,):
                raise ActionError("action must have shape (%d,)" % self.action_dim)
            a = a[None]
        if a.ndim == 2 and a.shape[1] == 4 and self.action_dim == 3:
            a = a[:, :3]  # BASELINE config "4-dim action" on a non-grasping task: the grip column is ignored
        if self.check_actions or a.shape != (self.batch, self.action_dim):
            self._check_action(a)
        self._np_action[...] = a
        # one C-ABI call: H2D of the actions, the step kernel, D2H of the four observation blocks (contiguous per
        # key, so handing them out is four memcpy-speed copies instead of strided de-interleaving) and the flags
        with torch.cuda.device(self.device):
            _lib.check(self._L.pmg_step_host_blocks(self._h, self._p_action, self._p_blocks, self._p_reward,
                                                    self._p_done, self._p_success, self._stream()))
        obs = {k: v.copy() for k, v in self._np_blocks.items()}
        if self._squeeze:
            obs = {k: v[0].astype(np.float64) for k, v in obs.items()}
        reward = self._np_reward.copy()
        done = self._np_done.view(np.bool_).copy()
        ok = self._np_success.view(np.bool_).copy()
        if not self.binary_reward:
            reward = reward.astype(np.float64)
        self.desired_goal = obs["desired_goal"]
        if self._squeeze:
            info = {"goal_achieved": bool(ok[0]), "is_success": bool(ok[0])}
            if done[0]:
                info["TimeLimit.truncated"] = True  # gym's TimeLimit sets the key only when the limit ends the episode
            return obs, reward[0], bool(done[0]), info
        # batched: the key is an array, True where the limit ended the episode
        info = {"goal_achieved": ok, "is_success": ok, "TimeLimit.truncated": done.copy()}
        if self._terminal is not None:
            info["terminal_observation"] = self._split(self._terminal)
        return obs, reward, done, info

    def _step_device(self, action):
        if action.dim() == 2 and action.shape[1] == 4 and self.action_dim == 3:
            action = action[:, :3]
        if tuple(action.shape) != (self.batch, self.action_dim):
            raise ActionError("action must have shape (%d, %d)" % (self.batch, self.action_dim))
        action = action.to(dtype=torch.float32).contiguous()
        if self.check_actions and self._L.pmg_action_error(self._h, 1):  # raised by the kernel of an EARLIER step (no sync here)
            raise ActionError("an earlier step was given an action outside the action space Box(-1, 1, (%d,))" % self.action_dim)
        with torch.cuda.device(self.device):
            out = torch.empty((self.batch, self.row_width), dtype=torch.float32, device=self.device)
            reward = torch.empty((self.batch,), dtype=torch.float32, device=self.device)
            flags = torch.empty((2, self.batch), dtype=torch.uint8, device=self.device)
            _lib.check(self._L.pmg_step(self._h, _ptr(action), _ptr(out), _ptr(reward), _ptr(flags[0]), _ptr(flags[1]), self._stream()))
        obs = self._split(out)
        self.desired_goal = obs["desired_goal"]
        done, ok = flags[0].bool(), flags[1].bool()
        info = {"goal_achieved": ok, "is_success": ok, "TimeLimit.truncated": done}
        if self._terminal is not None:
            info["terminal_observation"] = self._split(self._terminal)
        return obs, reward, done, info

    def action_error(self, synchronize=True):
        """True if a device-path step was given an action outside Box(-1, 1) (NaN included) since the last query;
        synchronises the device first by default, so that the answer covers every step enqueued so far."""
        if synchronize:
            torch.cuda.synchronize(self.device)
        return bool(self._L.pmg_action_error(self._h, 1))

    def step_packed(self, action, out, reward, done, success):
        """Zero-allocation device path: caller-owned CUDA buffers (used by bench.py and the sharded env)."""
        _lib.check(self._L.pmg_step(self._h, _ptr(action), _ptr(out), _ptr(reward), _ptr(done), _ptr(success), self._stream()))

    def _compute_reward(self, achieved_goal, desired_goal):
        """kuka_single_step_base_env.py:237-244: (reward, goal_achieved) with any leading axes."""
        if tuple(achieved_goal.shape) != tuple(desired_goal.shape):
            raise AssertionError("achieved_goal.shape != desired_goal.shape")
        as_numpy = not torch.is_tensor(achieved_goal)
        ag = torch.as_tensor(np.asarray(achieved_goal, dtype=np.float32) if as_numpy else achieved_goal).to(self.device, torch.float32).contiguous()
        dg = torch.as_tensor(np.asarray(desired_goal, dtype=np.float32) if as_numpy else desired_goal).to(self.device, torch.float32).contiguous()
        g = ag.shape[-1]
        lead = tuple(ag.shape[:-1])
        n = ag.numel() // g if g else 0
        with torch.cuda.device(self.device):
            r = torch.empty((n,), dtype=torch.float32, device=self.device)
            ok = torch.empty((n,), dtype=torch.uint8, device=self.device)
            _lib.check(self._L.pmg_compute_reward(_ptr(ag), _ptr(dg), n, g, self.distance_threshold, int(self.binary_reward),
                                                  _ptr(r), _ptr(ok), self._stream()))
        r, ok = r.reshape(lead), ok.bool().reshape(lead)
        if as_numpy:
            r, ok = r.cpu().numpy(), ok.cpu().numpy()
            if not self.binary_reward:
                r = r.astype(np.float64)
            if lead == ():
                return r[()], bool(ok)
        return r, ok

    def compute_reward(self, achieved_goal, desired_goal, info=None):
        """gym GoalEnv-style alias (not in the reference; BASELINE.json names it)."""
        return self._compute_reward(achieved_goal, desired_goal)[0]

    def render(self, mode="human", camera_id=0):
        raise NotImplementedError("rendering / image observations are outside the accelerated step path")

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._L.pmg_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- extras (no reference counterpart) ----------------------------------------------------
    def get_state(self):
        w = self._L.pmg_state_width(self._h)
        out = np.zeros((self.batch, w), dtype=np.float32)
        _lib.check(self._L.pmg_get_state(self._h, out.ctypes.data_as(C.c_void_p)))
        return out

    def set_state(self, state):
        w = self._L.pmg_state_width(self._h)
        s = np.ascontiguousarray(state, dtype=np.float32)
        if s.shape != (self.batch, w):
            raise ValueError("state must have shape (%d, %d)" % (self.batch, w))
        _lib.check(self._L.pmg_set_state(self._h, s.ctypes.data_as(C.c_void_p)))

    def last_spawn(self):
        out = np.zeros((self.batch, self._L.pmg_spawn_width(self._h)), dtype=np.float32)
        _lib.check(self._L.pmg_last_spawn(self._h, out.ctypes.data_as(C.c_void_p)))
        return out

    def kernel_timing(self, on=True):
        """Bracket every step kernel with CUDA events inside the library (bench.py's roofline measurement)."""
        _lib.check(self._L.pmg_kernel_timing(self._h, int(bool(on))))

    def kernel_time_ms(self):
        """-> (summed step-kernel milliseconds, launches covered) since kernel_timing(True); synchronises."""
        ms, n = C.c_double(), C.c_int64()
        _lib.check(self._L.pmg_kernel_time_ms(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    @property
    def launch_count(self):
        return int(self._L.pmg_launch_count(self._h))

    @property
    def overflow_count(self):
        return int(self._L.pmg_overflow_count(self._h))


class KukaReachEnv(KukaBulletMGEnv):  # kuka_single_step_envs.py:35-46
    def __init__(self, **kw):
        super().__init__("reach", **kw)


class KukaPushEnv(KukaBulletMGEnv):  # kuka_single_step_envs.py:20-32
    def __init__(self, **kw):
        super().__init__("push", **kw)


class KukaPickAndPlaceEnv(KukaBulletMGEnv):  # kuka_single_step_envs.py:4-17
    def __init__(self, **kw):
        super().__init__("pick_and_place", **kw)


class KukaSlideEnv(KukaBulletMGEnv):  # kuka_single_step_envs.py:49-59
    def __init__(self, **kw):
        super().__init__("slide", **kw)


class KukaBlockStackEnv(KukaBulletMGEnv):  # kuka_multi_step_envs.py:6-32
    def __init__(self, **kw):
        super().__init__("block_stack", **kw)


class KukaBlockRearrangeEnv(KukaBulletMGEnv):  # kuka_multi_step_envs.py:151-189
    def __init__(self, **kw):
        super().__init__("block_rearrange", **kw)
