/* pmg.h -- C-ABI of the B200-native batched Kuka multigoal simulator (libpmg.so).
 *
 * The reference (IanYangChina/pybullet_multigoal_gym) has no FFI layer of its own for this
 * path: `env.step()` is Python calling into the pybullet C extension.  The entry points
 * below are what a binding that replaces that path has to call; each one cites the reference
 * interface it stands in for (paths relative to /root/reference/pybullet_multigoal_gym/).
 * INTEGRATION.md shows the ctypes stub a maintainer of the reference would add.
 *
 * Conventions: every function returns 0 on success and a negative pmg_status on failure; the
 * message is available from pmg_last_error() (thread-local).  Nothing throws across the ABI.
 * `stream` arguments are cudaStream_t passed as void* (NULL = the legacy default stream).
 * `*_dev` pointers are device pointers on the handle's device, `*_host` are host pointers.
 * The caller owns every I/O buffer; the handle owns the persistent per-env state (SoA in HBM).
 * Calls on one handle must be serialised by the caller; different handles are independent.
 */
#ifndef PMG_H
#define PMG_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PMG_ABI_VERSION 6

typedef enum {
  PMG_OK = 0,
  PMG_ERR_INVALID = -1, /* bad argument (task name, num_block > 5, null pointer, ...) */
  PMG_ERR_CUDA = -2,    /* a CUDA runtime call failed; message holds cudaGetErrorString */
  PMG_ERR_STATE = -3    /* call order violated (e.g. step before the first reset) */
} pmg_status;

/* task ids: envs/task_envs/kuka_single_step_envs.py:4-46, kuka_multi_step_envs.py:6-32 (stack),
 * :151-189 (rearrange: the block-stack scene without grasping, one table target per block) */
typedef enum { PMG_REACH = 0, PMG_PUSH = 1, PMG_PICK_AND_PLACE = 2, PMG_BLOCK_STACK = 3, PMG_BLOCK_REARRANGE = 4,
               PMG_SLIDE = 5 /* kuka_single_step_envs.py:49-59: Push on the long low-friction table with a puck, goals beyond reach */ } pmg_task;

/* make_env(...) kwargs that reach the step path (__init__.py:4-11,88-131) + batch/device */
typedef struct {
  int32_t task;               /* pmg_task */
  int32_t num_block;          /* block_stack only, 1..5 (__init__.py:108) */
  int32_t batch;              /* environments stepped in lockstep on this handle */
  int32_t binary_reward;      /* 1: sparse -1/0 float32, 0: dense -distance */
  float distance_threshold;   /* default 0.05 (__init__.py:6) */
  int32_t max_episode_steps;  /* gym TimeLimit, default 50 (__init__.py:6,105) */
  int32_t device;             /* CUDA device ordinal */
  int32_t grip_informed_goal; /* block_stack only: goal += gripper xyz + finger closeness
                                 (kuka_multi_step_base_env.py:300-304, kuka_multi_step_envs.py:75-77) */
  int32_t joint_control;      /* 1: actions are 7 joint deltas of 0.05 rad (+ grip command) instead of a tip delta,
                                 joint poses are prepended to observation / policy_state (kuka.py:104-108,204-206;
                                 kuka_single_step_base_env.py:214-216) */
  int32_t task_decomposition; /* block_stack only: the desired goal is one of the sub-goals of
                                 kuka_multi_step_envs.py:88-120, selected with pmg_set_sub_goal */
  int32_t use_curriculum;     /* block_stack (exclusive with task_decomposition) and block_rearrange: every reset
                                 draws a goal difficulty level from a per-env probability schedule
                                 (kuka_multi_step_base_env.py:122-140,350-379; kuka_multi_step_envs.py:124-148);
                                 block_rearrange then gives targets to level + 1 randomly chosen blocks
                                 (kuka_multi_step_envs.py:193-227) */
  int32_t num_goals_to_generate; /* make_env(num_goals_to_generate=...): goals per level = this // num_block */
} pmg_config;

typedef struct pmg_handle pmg_handle;

int pmg_abi_version(void);
const char* pmg_last_error(void);

/* replaces: make_env() -> gym.make -> Kuka*Env.__init__ (__init__.py:178,
 * kuka_single_step_base_env.py:12-74, base_env.py:15-45,203-220).  Does NOT reset. */
int pmg_create(const pmg_config* cfg, pmg_handle** out);
int pmg_destroy(pmg_handle* h);

/* dims[0..5] = observation, policy_state, achieved_goal, desired_goal, action, packed-row width
 * (= sum of the first four).  replaces: observation_space / action_space (base_env.py:85-92). */
int pmg_dims(const pmg_handle* h, int32_t dims[6]);

/* replaces: env.seed() (base_env.py:120-122).  keys_host: [batch, max_key_len] uint32
 * MT19937 init_by_array keys (gym's sha512 seed hash is computed by the caller),
 * key_lens_host: [batch]. */
int pmg_seed(pmg_handle* h, const uint32_t* keys_host, const int32_t* key_lens_host, int32_t max_key_len);

/* replaces: env.reset() (base_env.py:124-128; kuka.py:157-165; kuka_single_step_base_env.py:
 * 104-148; kuka_multi_step_base_env.py:223-246; kuka_multi_step_envs.py:34-87).
 * mask_host: nullable [batch] bytes, non-zero = reset that env (NULL = all).
 * spawn_host: nullable [batch, spawn_width] floats = [block xy (2*nb) | desired_goal (G)]
 *   (G includes the 4 gripper entries of a grip-informed goal; curriculum handles append the sub-goal index
 *   equivalent to the drawn level: level, or 2 * level + 1 with grip-informed goals; block_rearrange: the bit
 *   mask of the blocks that have targets -- their goal words hold the targets, the other blocks' are ignored);
 *   NULL = sample on the host from each env's numpy-compatible MT19937 stream, exactly as the
 *   reference consumes it.
 * obs_dev: [batch, packed-row width] row-major, rows of envs not reset are rewritten unchanged. */
int pmg_reset(pmg_handle* h, const uint8_t* mask_host, const float* spawn_host, float* obs_dev, void* stream);
int pmg_spawn_width(const pmg_handle* h);
/* the spawn rows every env was last reset with, [batch, spawn_width] (for parity tests); after a device-sampled
 * reset (pmg_reset_device / auto-reset) the rows the reset kernel drew */
int pmg_last_spawn(const pmg_handle* h, float* spawn_host);

/* ---- device-side reset sampling and auto-reset ("spawn NULL => device Philox") ----------------
 * pmg_reset samples on ONE host thread from the reference's MT19937 streams (seed parity with the reference).  For
 * throughput runs the same sampling rules (kuka_single_step_base_env.py:104-148, kuka_multi_step_base_env.py:223-240,
 * kuka_multi_step_envs.py:34-87,174-189) are applied inside the reset kernel to a Philox4x32-10 stream per
 * (seed, env_index_base + env, episode number): no host work, no host<->device traffic, asynchronous on `stream`.
 * It is a different random stream from the reference's; oracle/device_rng_oracle.py restates it bit-exactly.
 * env_index_base: global index of this handle's env 0, so that a sharded batch draws the same rows whatever the
 * sharding.  Not available for curriculum handles (their schedule lives on the host). */
int pmg_set_device_rng(pmg_handle* h, uint64_t seed, int64_t env_index_base);
/* pmg_reset with device sampling; mask_dev: nullable [batch] bytes ON THE DEVICE (non-zero = reset that env). */
int pmg_reset_device(pmg_handle* h, const uint8_t* mask_dev, float* obs_dev, void* stream);
/* Auto-reset (gym VectorEnv semantics; the reference's single env leaves this to the caller's `if done: reset()`):
 * when on, every pmg_step / pmg_step_host* is followed on the same stream by a reset pass over the environments
 * whose `done` flag that step raised; their rows of obs_dev become the first observation of the new episode, while
 * reward / done / success stay those of the terminal step.  terminal_obs_dev: nullable [batch, W]; rows of the
 * environments that reset receive their terminal observation (the other rows are not written). */
int pmg_set_auto_reset(pmg_handle* h, int32_t on, float* terminal_obs_dev);

/* replaces: env.activate_curriculum_update() / env.deactivate_curriculum_update()
 * (kuka_multi_step_base_env.py:147-157); curriculum handles only. */
int pmg_set_curriculum_update(pmg_handle* h, int32_t on);
/* curriculum state of every env: prob_host [batch, num_block] (curriculum_prob), level_host [batch] (the level drawn
 * by the most recent reset; env.curriculum_goal_step = level * 25 + 50).  Either pointer may be NULL. */
int pmg_get_curriculum(const pmg_handle* h, float* prob_host, int32_t* level_host);

/* replaces: env.set_sub_goal(sub_goal_ind) (kuka_multi_step_base_env.py:159-181), task decomposition only.
 * ind_host: [batch] sub-goal index per env, python list indexing (-1 = the last sub-goal = the final goal; block
 * stack has num_block sub-goals, 2 * num_block with grip-informed goals: pick / place per level); NULL = -1 for
 * all.  The desired goal of the following observations is rebuilt from the current block positions
 * (kuka_multi_step_base_env.py:311-313); every reset selects -1 again (:247-248). */
int pmg_set_sub_goal(pmg_handle* h, const int32_t* ind_host, void* stream);

/* replaces: env.step(action) (base_env.py:130-138 -> kuka.py:167-225 -> 5 x stepSimulation;
 * kuka_single_step_base_env.py:193-244; kuka_multi_step_base_env.py:255-345; gym TimeLimit).
 * action_dev [batch, A] row-major in [-1, 1]; obs_dev [batch, W] packed
 * [observation | policy_state | achieved_goal | desired_goal]; reward_dev [batch];
 * done_dev / success_dev [batch] bytes (done = elapsed >= max_episode_steps,
 * success = info['goal_achieved']).  One kernel launch, asynchronous on `stream`. */
int pmg_step(pmg_handle* h, const float* action_dev, float* obs_dev, float* reward_dev,
             uint8_t* done_dev, uint8_t* success_dev, void* stream);

/* ---- multi-GPU: the batch sharded over ranks, the returned batch gathered over peer memory -------
 * One process per GPU, rank r owns environments [r * batch, (r + 1) * batch) of a global batch of world * batch.
 * Environments share nothing, so the only exchange is the returned batch.  Instead of an all-gather launched behind
 * the step kernel, the kernel's epilogue stores every finished environment's row, reward and flags straight into
 * slice r of EVERY rank's gather buffer (peer-mapped memory: the stores travel over NVLink / NVSwitch while the other
 * warps are still computing); the launch's last arrival publishes a per-rank sequence flag to all ranks and waits
 * for theirs, so the kernel's completion means "my buffer holds the global batch of this step".  (With auto-reset on,
 * the reset pass that follows the step kernel rewrites the rows of the finished environments and does the push.)
 *
 * pmg_gather_create allocates this rank's buffer and returns its 64-byte cudaIpcMemHandle; the caller exchanges the
 * handles of all ranks (any host channel, e.g. torch.distributed.all_gather_object) and passes the world x 64 bytes
 * to pmg_gather_connect.  world <= 8.  The buffer holds two copies (by step parity) of
 *   [obs f32 world*batch x W | reward f32 world*batch | done u8 world*batch | success u8 world*batch]
 * i.e. the global arrays, contiguous; pmg_gather_layout returns {bytes per parity copy, offsets of obs / reward /
 * done / success inside a copy, total bytes}.  pmg_step_gather returns the four device pointers of the copy this
 * step filled; they stay valid until the step after the next one (a peer may run at most one step ahead).
 * Every rank must call pmg_step_gather the same number of times. */
int pmg_gather_create(pmg_handle* h, int32_t rank, int32_t world, void* ipc_handle_out);
int pmg_gather_connect(pmg_handle* h, const void* ipc_handles);
int pmg_gather_layout(const pmg_handle* h, int64_t layout[6]);
int pmg_step_gather(pmg_handle* h, const float* action_dev, void** gathered_dev_out, void* stream);

/* Same call with HOST buffers (pinned or pageable): H2D of the actions, the step kernel, D2H of
 * obs / reward / done / success, then a stream synchronise.  This is the end-to-end entry a
 * reference-side binding calls once per env.step(). */
int pmg_step_host(pmg_handle* h, const float* action_host, float* obs_host, float* reward_host,
                  uint8_t* done_host, uint8_t* success_host, void* stream);

/* pmg_step_host with the observation delivered as four contiguous blocks instead of packed rows:
 * blocks_host = [observation [batch, O] | policy_state [batch, P] | achieved_goal [batch, G] | desired_goal [batch, G]],
 * each block row-major and contiguous -- the arrays of the reference's observation dict, ready to hand out without a
 * strided de-interleave on the host (which costs more than the PCIe transfer at batch 8192). */
int pmg_step_host_blocks(pmg_handle* h, const float* action_host, float* blocks_host, float* reward_host,
                         uint8_t* done_host, uint8_t* success_host, void* stream);

/* replaces: env._compute_reward(achieved_goal, desired_goal) on arbitrary leading axes
 * (kuka_single_step_base_env.py:237-244; kuka_multi_step_base_env.py:338-345), e.g. HER relabelling.
 * ag_dev, dg_dev: [n, g] row-major. */
int pmg_compute_reward(const float* ag_dev, const float* dg_dev, int64_t n, int32_t g, float threshold,
                       int32_t binary_reward, float* reward_dev, uint8_t* achieved_dev, void* stream);

/* ---- hindsight relabelling around _compute_reward ------------------------------------
 * The reference's agents (README.md:18-20) relabel stored transitions with goals achieved later in the same
 * episode ("future" strategy) and re-evaluate env._compute_reward on them.  Episodes live on the device:
 *   ag_dev [n_episodes, horizon + 1, g]  achieved goals, ag[e][t + 1] = achieved goal after transition t
 *   dg_dev [n_episodes, g]               the goals the episodes were collected with
 * pmg_her_sample draws n transitions: episode uniform, t uniform in [0, horizon), and with probability
 * her_prob a future index uniform in [t + 1, horizon] (else -1 = keep the original goal), from a counter-based
 * hash of (seed, sample index) -- oracle/her_oracle.py restates it bit-exactly in numpy.
 * pmg_her_relabel gathers the relabelled goal of every sample and evaluates _compute_reward(ag[e][t + 1], goal)
 * (kuka_single_step_base_env.py:237-244) in the same pass. */
int pmg_her_sample(int64_t n, int32_t n_episodes, int32_t horizon, float her_prob, uint64_t seed,
                   int32_t* episode_dev, int32_t* t_dev, int32_t* future_dev, void* stream);
int pmg_her_relabel(const float* ag_dev, const float* dg_dev, int32_t n_episodes, int32_t horizon, int32_t g,
                    const int32_t* episode_dev, const int32_t* t_dev, const int32_t* future_dev, int64_t n,
                    float threshold, int32_t binary_reward, float* goal_out_dev, float* reward_dev,
                    uint8_t* achieved_dev, void* stream);

/* State access for teacher-forced parity tests (no reference counterpart).  Row layout, floats:
 * q[9] qd[9] ee_target[3] rest_pose[7] motor_target[9] motor_max_impulse[9], then per block
 * pos[3] quat_xyzw[4] linvel[3] angvel[3], then desired_goal[G] (the final goal), then the sub-goal index (task
 * decomposition only), then elapsed steps.
 * pmg_set_state also clears the contact caches. */
int pmg_state_width(const pmg_handle* h);
int pmg_get_state(pmg_handle* h, float* state_host);
int pmg_set_state(pmg_handle* h, const float* state_host);

/* Measurement aid (bench.py's roofline): while on, every pmg_step / pmg_step_gather brackets its STEP KERNEL alone
 * (not the auto-reset pass, not the caller's copies) with CUDA events on the launch stream; pmg_kernel_time_ms
 * synchronises the device and returns the summed duration and the number of launches covered (a ring of the first
 * 2048 since the last pmg_kernel_timing call). */
int pmg_kernel_timing(pmg_handle* h, int32_t on);
int pmg_kernel_time_ms(pmg_handle* h, double* total_ms, int64_t* count);

/* Test aid (tests/test_gpu_physics.py): the box-box narrowphase of the path run on the device over n independent pairs,
 * one thread per pair.  in_host: n x 30 floats = p1[3] R1[9, row-major] half1[3] p2[3] R2[9] half2[3]; out_host: n x 32
 * floats = contact count, then <= 4 x (point on box 2 [3], normal on box 2 [3], signed distance).  stat: 0 = the
 * general 15-axis search, 1 / 2 = box 1 / box 2 is a static axis-aligned box (the fast path the step kernels use for
 * the table and the floor; results must be bit-identical to stat = 0). */
int pmg_debug_box_box(const float* in_host, int64_t n, int32_t stat, float* out_host, int32_t device);

/* replaces: `assert self.action_space.contains(a)` (kuka.py:168) on the DEVICE path, without a host round trip per step:
 * every step kernel tests its environments' action rows against Box(-1, 1) (NaN fails) and raises a word in mapped host
 * memory; the step itself still runs with the action as given.  Returns 1 if any step COMPLETED so far saw such an
 * action (the caller decides when to look: after a synchronise for an exact answer, or at the next call for an
 * asynchronous one, as CUDA reports its own errors); clear != 0 resets the word.  pmg_step_host* check their host
 * buffer before launching and return PMG_ERR_INVALID at once. */
int pmg_action_error(pmg_handle* h, int32_t clear);

/* number of kernels this library has launched on the handle since creation */
int64_t pmg_launch_count(const pmg_handle* h);
/* contact points dropped because a per-env scratch pool overflowed (0 in every shipped config) */
int64_t pmg_overflow_count(pmg_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* PMG_H */
