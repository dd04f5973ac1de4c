/* pmg_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, double precision, one environment at a time) of the hot path
 * `env.step()` of IanYangChina/pybullet_multigoal_gym for the Kuka iiwa14 + parallel-jaw
 * Reach / Push / PickAndPlace / BlockStack environments, including the part of the path
 * that lives in the un-vendored dependency pybullet ~= 3.0.6 (bullet3 C++).
 *
 * PARITY UNPINNED for the physics: pybullet/bullet3 are not available in the build
 * container or on the GPU box (no wheel, no sources, no network), and the reference's own
 * tests hold no assertions or golden vectors.  The Bullet
 * behaviours restated here are written from the published algorithm as recollected
 * (DESIGN.md "Bullet behaviours restated") and must be
 * confirmed with tools/dump_golden.py the first time a real pybullet is reachable.
 * What IS pinned: the reference's own Python plumbing (sampling, action map, observation
 * layout, reward) -- tests/golden/ holds vectors produced by executing the reference's
 * unmodified Python on top of this oracle through oracle/pybullet_shim -- and the RNG
 * (numpy's legacy RandomState, checked bit-exactly against numpy itself).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library.
 */
#ifndef PMG_ORACLE_H
#define PMG_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PMGO_MAX_BLOCKS 5
#define PMGO_MAX_OBS 176 /* nb=5, grip goal, joint control: 95 + 26 + 19 + 19 = 159 */

enum { PMGO_REACH = 0, PMGO_PUSH = 1, PMGO_PICK_AND_PLACE = 2, PMGO_BLOCK_STACK = 3, PMGO_BLOCK_REARRANGE = 4, PMGO_SLIDE = 5 };

typedef struct PmgoEnv PmgoEnv;

/* task: enum above; num_block used by BLOCK_STACK only (1..5). */
PmgoEnv* pmgo_create(int task, int num_block, int binary_reward, double distance_threshold,
                     int max_episode_steps);
/* + grip_informed_goal (block_stack: goal gains gripper xyz + finger closeness, kuka_multi_step_base_env.py:
 * 300-304, kuka_multi_step_envs.py:75-77) and joint_control (7 joint deltas instead of a tip delta,
 * kuka.py:104-108,204-206; joint poses prepended to observation / policy_state). */
PmgoEnv* pmgo_create_ex(int task, int num_block, int binary_reward, double distance_threshold,
                        int max_episode_steps, int grip_informed_goal, int joint_control);
/* + task_decomposition (block_stack: the desired goal is one of the sub-goals of kuka_multi_step_envs.py:88-120,
 * chosen with pmgo_set_sub_goal like env.set_sub_goal, kuka_multi_step_base_env.py:159-165; reset selects -1). */
PmgoEnv* pmgo_create_ex2(int task, int num_block, int binary_reward, double distance_threshold,
                         int max_episode_steps, int grip_informed_goal, int joint_control, int task_decomposition);
/* + use_curriculum (block_stack; kuka_multi_step_base_env.py:122-140,350-379, kuka_multi_step_envs.py:124-148): every
 * reset draws a difficulty level from curriculum_prob (np_random.choice) and only the stack levels <= it have to
 * be built; with updates activated the probabilities follow the number of generated goals per level. */
PmgoEnv* pmgo_create_ex3(int task, int num_block, int binary_reward, double distance_threshold, int max_episode_steps,
                         int grip_informed_goal, int joint_control, int task_decomposition, int use_curriculum,
                         long num_goals_to_generate);
void pmgo_set_curriculum_update(PmgoEnv* e, int on); /* activate_ / deactivate_curriculum_update */
int pmgo_get_curriculum(const PmgoEnv* e, double* prob_out /*[num_block]*/); /* returns the current level */
void pmgo_set_sub_goal(PmgoEnv* e, int sub_goal_ind);
/* the observation of the current state, without stepping (_get_obs) */
void pmgo_observe(PmgoEnv* e, double* obs_out);
void pmgo_destroy(PmgoEnv* e);

/* dims[0..3] = observation, policy_state, achieved_goal, desired_goal lengths; returns action dim */
int pmgo_dims(const PmgoEnv* e, int dims[4]);

/* numpy RandomState.seed(list-of-uint32) == MT19937 init_by_array (gym seeding hashes the
 * integer seed with sha512 on the Python side, base_env.py:120-122). */
void pmgo_seed_array(PmgoEnv* e, const uint32_t* key, int len);

/* reset following base_env.py:124-128; samples object / goal poses from the env's own
 * MT19937 stream exactly as kuka_single_step_base_env.py:104-148 /
 * kuka_multi_step_base_env.py:223-246 / kuka_multi_step_envs.py:34-87 do.
 * obs_out: packed [observation | policy_state | achieved_goal | desired_goal]. */
void pmgo_reset(PmgoEnv* e, double* obs_out);

/* reset with caller-provided spawn data instead of the RNG:
 * spawn = [block xy (2*nb) | desired_goal (G)] ; used to mirror a GPU reset. */
void pmgo_reset_with(PmgoEnv* e, const double* spawn, double* obs_out);

/* one env.step(action) (base_env.py:130-138 + gym TimeLimit). reward is the float64 value the
 * reference would return (sparse: -0.0 / -1.0). */
void pmgo_step(PmgoEnv* e, const double* action, double* obs_out, double* reward, int* done,
               int* goal_achieved);

/* _compute_reward on arbitrary rows (HER relabelling): ag, dg [n, g] row-major. */
void pmgo_compute_reward(const double* ag, const double* dg, int64_t n, int g, double thr,
                         int binary, double* reward, uint8_t* achieved);

/* ---- state access for teacher-forced parity tests -------------------------------------- */
/* layout: q[9] qd[9] ee_target[3] rest_pose[7] motor_target[9] motor_maximp[9]
 *         then per block pos[3] quat[4](xyzw) linvel[3] angvel[3]; then desired_goal[G] (the final goal);
 *         then sub_goal_ind (task decomposition / curriculum only); then elapsed (as double). */
int pmgo_state_size(const PmgoEnv* e);
void pmgo_get_state(const PmgoEnv* e, double* out);
void pmgo_set_state(PmgoEnv* e, const double* in); /* also clears the contact caches */
void pmgo_poke_state(PmgoEnv* e, const double* in); /* same, contact caches kept (pybullet_shim) */

/* ---- pieces exposed for unit tests and for oracle/pybullet_shim ------------------------ */
void pmgo_fk_tip(const double q[9], double pos[3], double quat_xyzw[4]);
/* calculateInverseKinematics restatement: seed q (9), target pos, quat; returns 9 values. */
void pmgo_ik(const double q_seed[9], const double target_pos[3], const double target_quat[4],
             int max_iter, double residual_threshold, double q_out[9]);
/* advance the physics by n substeps of 0.002 s with the current motor settings. */
void pmgo_substeps(PmgoEnv* e, int n);
/* stepSimulation(): joint-damping torque sample + 20 substeps. */
void pmgo_step_simulation(PmgoEnv* e);
/* mass matrix via unit-impulse responses of the ABA (for cross-checks): out[81] row-major */
void pmgo_mass_matrix_inverse(PmgoEnv* e, double* out81);
/* contact dump: up to max rows of [pair, pointA(3), pointB(3), normalB(3), distance]; returns count */
int pmgo_get_contacts(const PmgoEnv* e, double* out, int max);
/* link world state like getLinkState(computeLinkVelocity=1) for: 0 tip, 1 gripper base,
 * 2 finger1 tab, 3 finger2 tab, 4 finger1, 5 finger2 ; out = pos3 quat4 linvel3 angvel3 */
void pmgo_link_state(const PmgoEnv* e, int which, double out[13]);

/* numpy legacy RandomState restatement, exposed for the bit-exact RNG tests */
void pmgo_rng_uniform(PmgoEnv* e, double lo, double hi, int n, double* out);
void pmgo_rng_shuffle(PmgoEnv* e, int64_t* arr, int n);

/* batch driver for the CPU baseline: steps `n_env` independent envs `n_steps` times with
 * actions[n_steps][n_env][adim] using `n_threads` pthreads; returns wall seconds. */
double pmgo_bench_rollout(PmgoEnv** envs, int n_env, const double* actions, int n_steps,
                          int n_threads);

#ifdef __cplusplus
}
#endif
#endif
