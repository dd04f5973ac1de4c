// pmg_capi.cu -- kernels and the C-ABI (include/pmg.h) of the batched Kuka multigoal simulator.
//
// Kernels (one thread = one environment, see pmg_physics.cuh for why):
//   step_kernel<TASK,NBLK>   : action map + IK + 5 x 20 substeps + observation/reward/flags
//   reset_kernel<TASK,NBLK>  : robot reset (IK to the start pose), object/goal placement, observation
//   reward_kernel            : _compute_reward over arbitrary rows (HER relabelling)
// Host side: handle bookkeeping and the numpy-compatible MT19937 reset sampler.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <vector>

#include "../../include/pmg.h"
#include "pmg_sim.cuh"
#include "pmg_coop.cuh"
#include "pmg_spawn.cuh"

using namespace pmg;

// ------------------------------------------------------------------------------------------------
// device: observation assembly (kuka.py:227-256, kuka_single_step_base_env.py:193-221,
// kuka_multi_step_base_env.py:255-320) and reward (kuka_single_step_base_env.py:237-244)
// ------------------------------------------------------------------------------------------------
// env index owned by this thread, or -1 for an idle lane / past the end of the batch
__device__ __forceinline__ int env_of_thread(const StepIO& io) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int i = warp * io.epw + lane;
  return (lane < io.epw && i < io.batch) ? i : -1;
}

__device__ __forceinline__ float clip5(float v) { return fminf(fmaxf(v, -5.0f), 5.0f); }

// ---- TMA engine: 1-D bulk async copies global -> shared, completion on an mbarrier -------------
// The persistent state is [word][env]; the 32 environments of a warp are one 128-byte row per word.
// Every lane issues the bulk copies of its share of the rows into a [word][32] shared-memory tile and
// the warp waits once on the mbarrier, instead of ~50-130 dependent scalar global loads per thread.
// Measured (profiles/r01_bulk_copy_carveout_ab.txt): correct, but the step is 25-30 % SLOWER with it --
// the prologue is microseconds of a multi-millisecond latency-bound kernel and the tile costs L1
// capacity that the per-thread scratch needs -- so it is opt-in (PMG_BULK_COPY=1), off by default.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int WORDS>
__device__ __forceinline__ void bulk_load_state_tile(float* tile, const float* gsrc, size_t batch, uint64_t* bar) {
  const int lane = threadIdx.x & 31;
  const uint32_t bar_a = smem_u32(bar);
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_a), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(WORDS * 128) : "memory");
  }
  __syncwarp();
  for (int w = lane; w < WORDS; w += 32)
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(tile + w * 32)), "l"(gsrc + (size_t)w * batch), "r"(128), "r"(bar_a) : "memory");
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar_a), "r"(0) : "memory");
}

// s points at this environment's word 0, consecutive words are `B` floats apart
// (global state: s = state + env, B = batch; staged tile: s = tile + lane, B = 32).
template <int TASK, int NBLK>
__device__ void load_env(Env<NBLK>& e, const StepIO& io, int i, const float* s, const size_t B) {
#pragma unroll
  for (int k = 0; k < ND; k++) {
    e.q[k] = s[(ST_Q + k) * B]; e.qd[k] = s[(ST_QD + k) * B];
    e.mt[k] = s[(ST_MT + k) * B]; e.mi[k] = s[(ST_MI + k) * B]; e.dtau[k] = 0.0f;
  }
#pragma unroll
  for (int b = 0; b < NBLK; b++) {
    const float* bs = s + (ST_BLK + 13 * b) * B;
    e.bpos[b] = v3(bs[0], bs[B], bs[2 * B]);
#pragma unroll
    for (int k = 0; k < 4; k++) e.bquat[b][k] = bs[(3 + k) * B];
    e.bv[b] = v3(bs[7 * B], bs[8 * B], bs[9 * B]);
    e.bw[b] = v3(bs[10 * B], bs[11 * B], bs[12 * B]);
  }
  e.man = io.manifold + man_off(io, i); e.stride = io.tile; e.overflow = 0;
}

template <int TASK, int NBLK>
__device__ void store_env(const Env<NBLK>& e, const StepIO& io, int i) {
  const size_t B = io.tile;  // distance between consecutive words of an environment (StepIO::tile)
  float* s = io.state + state_off(io, i);
#pragma unroll
  for (int k = 0; k < ND; k++) {
    s[(ST_Q + k) * B] = e.q[k]; s[(ST_QD + k) * B] = e.qd[k];
    s[(ST_MT + k) * B] = e.mt[k]; s[(ST_MI + k) * B] = e.mi[k];
  }
#pragma unroll
  for (int b = 0; b < NBLK; b++) {
    float* bs = s + (ST_BLK + 13 * b) * B;
    bs[0] = e.bpos[b].x; bs[B] = e.bpos[b].y; bs[2 * B] = e.bpos[b].z;
#pragma unroll
    for (int k = 0; k < 4; k++) bs[(3 + k) * B] = e.bquat[b][k];
    bs[7 * B] = e.bv[b].x; bs[8 * B] = e.bv[b].y; bs[9 * B] = e.bv[b].z;
    bs[10 * B] = e.bw[b].x; bs[11 * B] = e.bw[b].y; bs[12 * B] = e.bw[b].z;
  }
  if (e.overflow) atomicAdd(io.overflow, e.overflow);
}

// Stage the warp's rows in shared memory so that the global stores are contiguous 128-byte lines
// (rows of consecutive envs are adjacent in the packed [batch, W] output).  Every lane of the warp
// must call this; lanes that own no environment pass live = false.
__device__ __forceinline__ void stage_row(const float* row, const StepIO& io, bool live) {
  extern __shared__ float stage[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int W = io.row_width;
  float* ws = stage + warp * io.epw * W;
  if (live) {
    for (int k = 0; k < W; k++) ws[lane * W + k] = row[k];
  }
  __syncwarp();
  const int env0 = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * io.epw;
  const int nvalid = max(0, min(io.epw, io.batch - env0)) * W;
  float* out = io.obs + (size_t)env0 * W;
  for (int k = lane; k < nvalid; k += 32) out[k] = ws[k];
  if (io.g_n && io.g_in_step)  // fused gather: the same lines go to this rank's slice of every peer's buffer
    for (int d = 0; d < io.g_n; d++) {
      float* rout = io.g_obs[d] + (size_t)env0 * W;
      for (int k = lane; k < nvalid; k += 32) rout[k] = ws[k];
    }
  if (io.g_n && io.g_in_step) __threadfence_system();  // idle lanes store too but never arrive: order their stores here
  __syncwarp();
}

// thread-per-env kernels: reward and flags of environment i to the peers, then sign it off
__device__ __forceinline__ void gather_finish_thread(const StepIO& io, int i) {
  if (io.g_n == 0 || !io.g_in_step) return;
  for (int d = 0; d < io.g_n; d++) { io.g_reward[d][i] = io.reward[i]; io.g_done[d][i] = io.done[i]; io.g_success[d][i] = io.success[i]; }
  gather_arrive(io);
}

// Writes the packed row [observation | policy_state | achieved_goal | desired_goal] and returns
// the goal distance.  Run-time variants: joint control prepends the 7 arm joint angles to observation and
// policy_state, a grip-informed goal appends gripper xyz + finger closeness to the achieved goal.
// DIRECT: the thread stores its row itself (the auto-reset path, where only some lanes of a warp have a row to write).
template <int TASK, int NBLK, bool DIRECT = false>
__device__ float write_obs(const Env<NBLK>& e, const StepIO& io, int i) {
  using D = Dims<TASK, NBLK>;
  const size_t B = io.tile;
  Frames f;
  forward_kinematics<NB>(e.q, f);
  V3 tip = tip_position(f), tv, tw;
  point_velocity<PMG_BODY_LINK7>(f, e.qd, tip, tv, tw);
  float closeness = 0.0f, finger_vel = 0.0f;
  if (TASK >= 2 && io.grasp) {
    const float t1[3] = PMG_TAB1_OFFSET, t2[3] = PMG_TAB2_OFFSET;
    V3 tab1 = f.p[PMG_BODY_FINGER1] + mul(f.R[PMG_BODY_FINGER1], v3(t1[0], t1[1], t1[2]));
    V3 tab2 = f.p[PMG_BODY_FINGER2] + mul(f.R[PMG_BODY_FINGER2], v3(t2[0], t2[1], t2[2]));
    closeness = norm(tab1 - tab2);
    V3 vb, wb, vt, wt;
    point_velocity<PMG_BODY_GBASE>(f, e.qd, f.p[PMG_BODY_GBASE], vb, wb);
    point_velocity<PMG_BODY_FINGER1>(f, e.qd, tab1, vt, wt);
    finger_vel = vb.y - vt.y;
  }
  const int jo = io.jc ? 7 : 0, G = io.goal_dim;
  const int O = D::O + jo, P = D::P + jo;
  float row[D::W + 14 + 8];
  float* obs = row + jo; float* pol = row + O + jo; float* ag = row + O + P; float* dg = ag + G;
  if (io.jc) {
#pragma unroll
    for (int k = 0; k < 7; k++) { row[k] = e.q[k]; row[O + k] = e.q[k]; }
  }
  const float* goal = io.state + state_off(io, i) + (size_t)(ST_BLK + 13 * NBLK) * B;
  for (int k = 0; k < G; k++) dg[k] = goal[k * B];
  if (TASK == 0) {
    obs[0] = pol[0] = ag[0] = tip.x; obs[1] = pol[1] = ag[1] = tip.y; obs[2] = pol[2] = ag[2] = tip.z;
  } else if (TASK != 3) {
    V3 bx = e.bpos[0], rel = tip - bx, rv = tv - e.bv[0], rw = tw - e.bw[0];
    obs[0] = tip.x; obs[1] = tip.y; obs[2] = tip.z; obs[3] = bx.x; obs[4] = bx.y; obs[5] = bx.z; obs[6] = closeness;
    obs[7] = rel.x; obs[8] = rel.y; obs[9] = rel.z; obs[10] = tv.x; obs[11] = tv.y; obs[12] = tv.z; obs[13] = finger_vel;
    obs[14] = rv.x; obs[15] = rv.y; obs[16] = rv.z; obs[17] = rw.x; obs[18] = rw.y; obs[19] = rw.z;
    pol[0] = tip.x; pol[1] = tip.y; pol[2] = tip.z; pol[3] = closeness; pol[4] = rel.x; pol[5] = rel.y; pol[6] = rel.z;
    ag[0] = bx.x; ag[1] = bx.y; ag[2] = bx.z;
  } else {
    obs[0] = tip.x; obs[1] = tip.y; obs[2] = tip.z; obs[3] = closeness; obs[4] = tv.x; obs[5] = tv.y; obs[6] = tv.z; obs[7] = finger_vel;
    pol[0] = tip.x; pol[1] = tip.y; pol[2] = tip.z; pol[3] = closeness;
#pragma unroll
    for (int n = 0; n < NBLK; n++) {
      float* bs = obs + 8 + 16 * n;
      V3 bx = e.bpos[n], rel = tip - bx, rv = tv - e.bv[n], rw = tw - e.bw[n];
      bs[0] = bx.x; bs[1] = bx.y; bs[2] = bx.z; bs[3] = rel.x; bs[4] = rel.y; bs[5] = rel.z;
      bs[6] = e.bquat[n][0]; bs[7] = e.bquat[n][1]; bs[8] = e.bquat[n][2]; bs[9] = e.bquat[n][3];
      bs[10] = rv.x; bs[11] = rv.y; bs[12] = rv.z; bs[13] = rw.x; bs[14] = rw.y; bs[15] = rw.z;
      pol[4 + 3 * n] = rel.x; pol[5 + 3 * n] = rel.y; pol[6 + 3 * n] = rel.z;
      ag[3 * n] = bx.x; ag[3 * n + 1] = bx.y; ag[3 * n + 2] = bx.z;
    }
    if (io.grip_goal) { ag[3 * NBLK] = tip.x; ag[3 * NBLK + 1] = tip.y; ag[3 * NBLK + 2] = tip.z; ag[3 * NBLK + 3] = closeness; }
    if (io.td == 2) {
      // BlockRearrange curriculum (kuka_multi_step_envs.py:193-227): the state word behind the goal is the bit mask of
      // the blocks this episode moves; their goal words hold their targets, every other block's goal is where it is
      const int moved = (int)goal[(size_t)G * B];
#pragma unroll
      for (int n = 0; n < NBLK; n++)
        if (!((moved >> n) & 1)) { dg[3 * n] = e.bpos[n].x; dg[3 * n + 1] = e.bpos[n].y; dg[3 * n + 2] = e.bpos[n].z; }
    } else if (io.td) {
      // Task decomposition (kuka_multi_step_envs.py:88-120, kuka_multi_step_base_env.py:159-165,311-313): the
      // desired goal is sub_goals[ind], rebuilt from the current block positions.  The stored goal is the final
      // one; a block's stack level is its target height.  Without the grip goal sub-goal k has levels <= k on
      // their targets and the other blocks where they are; with it there is a pick (2k) / place (2k+1) pair.
      const int nsub = io.grip_goal ? 2 * NBLK : NBLK;
      int ind = (int)goal[(size_t)G * B];
      if (ind < 0) ind += nsub;
      const int k = io.grip_goal ? ind >> 1 : ind;
      const bool place = io.grip_goal ? (ind & 1) != 0 : true;
#pragma unroll
      for (int n = 0; n < NBLK; n++) {
        const int level = (int)floorf((dg[3 * n + 2] - BLOCK_SPAWN_Z) * (1.0f / 0.03f) + 0.5f);
        const bool at_target = place ? level <= k : level < k;
        if (io.grip_goal && level == k) {
          dg[3 * NBLK] = place ? dg[3 * n] : e.bpos[n].x; dg[3 * NBLK + 1] = place ? dg[3 * n + 1] : e.bpos[n].y;
          dg[3 * NBLK + 2] = place ? dg[3 * n + 2] : e.bpos[n].z;
        }
        if (!at_target) { dg[3 * n] = e.bpos[n].x; dg[3 * n + 1] = e.bpos[n].y; dg[3 * n + 2] = e.bpos[n].z; }
      }
    }
    for (int k = 0; k < O + P; k++) row[k] = clip5(row[k]);  // np.clip over the concatenated vectors, joint poses included
  }
  float d2 = 0.0f;
  for (int k = 0; k < G; k++) { float d = ag[k] - dg[k]; d2 += d * d; }
  if (DIRECT) {
    float* out = io.obs + (size_t)i * io.row_width;
    for (int k = 0; k < io.row_width; k++) out[k] = row[k];
  } else {
    stage_row(row, io, true);
  }
  return sqrtf(d2);
}

template <int TASK, int NBLK>
__global__ void __launch_bounds__(32) step_kernel(StepIO io) {
  using D = Dims<TASK, NBLK>;
  const int i = env_of_thread(io);
  const size_t B = io.tile;
  if (i >= 0) check_action_row(io, i, 0, 1);
  Env<NBLK> e;
  float ee0[3];
  if (io.bulk) {  // full warps only (epw == 32, batch % 32 == 0): no idle lanes on this path
    extern __shared__ float dyn_smem[];
    __shared__ __align__(8) uint64_t bar;
    float* tile = dyn_smem + io.tile_offset;
    const int env0 = (blockIdx.x * blockDim.x + threadIdx.x) & ~31;
    bulk_load_state_tile<D::STATE>(tile, io.state + state_off(io, env0), B, &bar);
    const float* ts = tile + (threadIdx.x & 31);
    load_env<TASK, NBLK>(e, io, i, ts, 32);
#pragma unroll
    for (int k = 0; k < 3; k++) ee0[k] = ts[(ST_EE + k) * 32];
  } else {
    if (i < 0) { stage_row(nullptr, io, false); return; }  // idle lanes only help the staged store
    load_env<TASK, NBLK>(e, io, i, io.state + state_off(io, i), B);
#pragma unroll
    for (int k = 0; k < 3; k++) ee0[k] = io.state[state_off(io, i) + (ST_EE + k) * B];
  }
  float* s = io.state + state_off(io, max(i, 0));
  // ---- Kuka.apply_action (kuka.py:167-222) ----
  float a[8];
#pragma unroll
  for (int k = 0; k < 8; k++) a[k] = k < io.adim ? io.action[(size_t)i * io.adim + k] : 0.0f;
  if (TASK >= 2 && io.grasp) {
    float grip = ((io.jc ? a[7] : a[3]) + 1.0f) * (GRIPPER_ABS_LIMIT / 2);
    e.mt[7] = e.mt[8] = grip; e.mi[7] = e.mi[8] = FINGER_FORCE * OUTER_DT;
  }
  const float lo[3] = {-0.67f, -0.20f, 0.175f}, hi[3] = {-0.37f, 0.20f, 0.55f};  // kuka.py:40-41
  float ee[3];
  if (io.jc) {
    // kuka.py:204-206: joint_state_target (kept in the motor targets) += 0.05 a[:7]; no clipping, no IK
#pragma unroll
    for (int k = 0; k < 3; k++) ee[k] = ee0[k];
#pragma unroll
    for (int k = 0; k < 7; k++) { e.mt[k] += a[k] * 0.05f; e.mi[k] = ARM_FORCE * OUTER_DT; }
  } else {
#pragma unroll
  for (int k = 0; k < 3; k++) ee[k] = fminf(fmaxf(ee0[k] + a[k] * 0.01f, lo[k]), hi[k]);
  {
    float qik[ND];
#pragma unroll
    for (int k = 0; k < ND; k++) qik[k] = e.q[k];
    const float tq[4] = {0.f, -1.f, 0.f, 0.f};  // kuka.py:42
    inverse_kinematics(qik, v3(ee[0], ee[1], ee[2]), tq);
#pragma unroll
    for (int k = 0; k < 7; k++) { e.mt[k] = qik[k]; e.mi[k] = ARM_FORCE * OUTER_DT; }
  }
  }
  // ---- 5 x stepSimulation (kuka.py:223-225), each 20 substeps of 2 ms ----
  for (int call = 0; call < CALLS_PER_ENV_STEP; call++) {
#pragma unroll
    for (int k = 0; k < ND; k++) e.dtau[k] = -c_dof_damping[k] * e.qd[k];
    for (int sub = 0; sub < SUBSTEPS_PER_CALL; sub++) substep(e);
  }
  float dist = write_obs<TASK, NBLK>(e, io, i);
  store_env<TASK, NBLK>(e, io, i);
#pragma unroll
  for (int k = 0; k < 3; k++) s[(ST_EE + k) * B] = ee[k];
  float* el = s + (size_t)(io.state_words - 1) * B;
  int elapsed = (int)(*el) + 1;
  *el = (float)elapsed;
  bool na = dist > io.thr;
  io.reward[i] = io.binary ? -(na ? 1.0f : 0.0f) : -dist;
  io.success[i] = na ? 0 : 1;
  io.done[i] = elapsed >= io.max_steps ? 1 : 0;
  gather_finish_thread(io, i);
}

// Lane-cooperative Reach step (pmg_coop.cuh): 8 lanes per environment, 4 environments per one-warp block.
// (Packing the environments whose jaws rest on the table into the same warps was tried and measured slower:
// the step lasts as long as its slowest warp, and four contact octets in one warp serialise their divergent
// narrowphase branches.)
constexpr int COOP_TABLE_BYTES = coop::GL * coop::LC_W * sizeof(float);
static_assert(COOP_TABLE_BYTES % 16 == 0, "EnvSmem must stay 16-byte aligned behind the constant table");

// The per-body model constants of the chain (joint frames, centres of mass, inertias, masses, limits, damping: 8 rows of
// LC_W floats, pmg_coop.cuh LC_*) live once per device in global memory, laid out exactly as the kernels read them, and
// every block stages them into its shared memory with ONE bulk asynchronous copy (TMA, cp.async.bulk, 896 bytes) that
// completes on an mbarrier -- instead of eight threads assembling the table from ~30 __constant__ loads each.
__device__ __align__(16) float g_lane_table[coop::GL * coop::LC_W];
__global__ void fill_lane_table_kernel() {
  if (threadIdx.x < coop::GL) coop::fill_lane_constants(g_lane_table + threadIdx.x * coop::LC_W, threadIdx.x);
}
// Every thread of the block calls this; returns when the table is in `dst` (16-byte aligned shared memory).
__device__ __forceinline__ void stage_lane_table(float* dst) {
  __shared__ __align__(8) uint64_t table_bar;
  const uint32_t bar_a = smem_u32(&table_bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_a), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(COOP_TABLE_BYTES) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(g_lane_table), "r"(COOP_TABLE_BYTES), "r"(bar_a) : "memory");
  }
  if (blockDim.x > 32) __syncthreads(); else __syncwarp();  // the barrier is initialised before anybody polls it
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar_a), "r"(0) : "memory");
}
constexpr int COOP_MIN_BLOCKS = 14;  // batch 8192 = 2048 blocks = 13.8 per SM: keep them all resident
template <bool JC>
__global__ void __launch_bounds__(64, COOP_MIN_BLOCKS / 2) step_kernel_coop_reach(StepIO io) {
  extern __shared__ __align__(16) unsigned char coop_smem[];
  const int lane32 = threadIdx.x & 31, grp = lane32 >> 3, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  float* lane_consts = reinterpret_cast<float*>(coop_smem);
  stage_lane_table(lane_consts);
  const int env = (blockIdx.x * wpb + warp) * io.epb + grp;  // io.epb environments per warp (octets beyond it stay idle)
  if (grp >= io.epb || env >= io.batch) return;   // a whole octet leaves together (exited threads do not hold up the block barrier)
  coop::Grp g;
  g.lane = lane32 & (coop::GL - 1); g.shift = grp * coop::GL; g.mask = 0xffu << g.shift;
  coop::EnvSmem& sm = *reinterpret_cast<coop::EnvSmem*>(coop_smem + COOP_TABLE_BYTES + (warp * io.epb + grp) * coop::env_stride<coop::EnvSmem>());
  check_action_row(io, env, g.lane, coop::GL);
  coop::step_env_reach<JC>(g, sm, lane_consts, io, env);
}

// Lane-cooperative Push / PickAndPlace step: the same octet layout with the block and its manifolds in shared memory
// (7.1 KB per environment with the first 12 contact points' rows; 7 one-warp blocks = 28 environments per SM, so a
// 4096-environment batch is one wave of 1024 blocks on 148 SMs).
template <int TASK>
__global__ void __launch_bounds__(64, 3) step_kernel_coop_block(StepIO io) {
  extern __shared__ __align__(16) unsigned char coop_smem[];
  const int lane32 = threadIdx.x & 31, grp = lane32 >> 3, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  float* lane_consts = reinterpret_cast<float*>(coop_smem);
  stage_lane_table(lane_consts);
  const int env = (blockIdx.x * wpb + warp) * io.epb + grp;  // io.epb environments per warp (octets beyond it stay idle)
  if (grp >= io.epb || env >= io.batch) return;   // a whole octet leaves together (exited threads do not hold up the block barrier)
  coop::Grp g;
  g.lane = lane32 & (coop::GL - 1); g.shift = grp * coop::GL; g.mask = 0xffu << g.shift;
  coop::EnvSmemT<1, TASK == 5>& sm = *reinterpret_cast<coop::EnvSmemT<1, TASK == 5>*>(coop_smem + COOP_TABLE_BYTES + (warp * io.epb + grp) * coop::env_stride<coop::EnvSmemT<1, TASK == 5>>());
  check_action_row(io, env, g.lane, coop::GL);
  coop::step_env_block<TASK>(g, sm, lane_consts, io, env);
}

// Lane-cooperative BlockStack / BlockRearrange step (NBLK = 2..5), the default for these tasks since round 2: it passes
// the same GPU parity tests as the thread-per-env kernel (PMG_COOP_STACK=0), racecheck-clean, 1.9x its rate at B = 2048.
template <int NBLK>
__global__ void __launch_bounds__(128, 1) step_kernel_coop_multi(StepIO io) {
  extern __shared__ __align__(16) unsigned char coop_smem[];
  const int lane32 = threadIdx.x & 31, grp = lane32 >> 3, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  float* lane_consts = reinterpret_cast<float*>(coop_smem);
  stage_lane_table(lane_consts);
  const int env = (blockIdx.x * wpb + warp) * io.epb + grp;  // io.epb environments per warp (octets beyond it stay idle)
  if (grp >= io.epb || env >= io.batch) return;   // a whole octet leaves together (exited threads do not hold up the block barrier)
  coop::Grp g;
  g.lane = lane32 & (coop::GL - 1); g.shift = grp * coop::GL; g.mask = 0xffu << g.shift;
  coop::EnvSmemT<NBLK>& sm = *reinterpret_cast<coop::EnvSmemT<NBLK>*>(coop_smem + COOP_TABLE_BYTES + (warp * io.epb + grp) * coop::env_stride<coop::EnvSmemT<NBLK>>());
  check_action_row(io, env, g.lane, coop::GL);
  coop::step_env_multi<NBLK>(g, sm, lane_consts, io, env);
}

// mask / spawn are device pointers.  spawn == nullptr: the row is sampled here from the environment's Philox stream
// (pmg_spawn.cuh) and recorded in spawn_out.  auto_rows: the auto-reset pass behind a step -- mask is that step's
// `done` flags, only the rows of the environments that reset are rewritten (with their first observation of the new
// episode; the terminal row is copied to `terminal` first when given), reward / done / success stay the terminal ones.
struct ResetIO {
  StepIO io; const uint8_t* mask; const float* spawn; float tip_init[3];
  float spawn_z;  // height the objects are placed at: cubes 0.175, Slide's puck 0.170
  int task, auto_rows; unsigned long long seed; long long env_base; uint32_t* episode; float* spawn_out; float* terminal;
  spawn::Bounds bounds;
};

// Reset (doit) and observation of environment i.  Auto-reset pass: called for the finished environments only, the row
// is stored by the thread itself and the terminal row is saved first when asked for.
template <int TASK, int NBLK>
__device__ void reset_env(const ResetIO& r, int i, bool doit) {
  const StepIO& io = r.io;
  const size_t B = io.tile;
  if (r.auto_rows && r.terminal) {
    const float* src = io.obs + (size_t)i * io.row_width;
    float* dst = r.terminal + (size_t)i * io.row_width;
    for (int k = 0; k < io.row_width; k++) dst[k] = src[k];
  }

  Env<NBLK> e;
  load_env<TASK, NBLK>(e, io, i, io.state + state_off(io, i), B);
  float* s = io.state + state_off(io, i);
  if (doit) {
    // Kuka.robot_specific_reset (kuka.py:157-165): joints to the rest pose, rest pose <- IK(start
    // position) seeded there, joints to the new rest pose, jaws closed with their motor on.
    float qik[ND];
#pragma unroll
    for (int k = 0; k < 7; k++) qik[k] = s[(ST_REST + k) * B];
    qik[7] = e.q[7]; qik[8] = e.q[8];
    const float tq[4] = {0.f, -1.f, 0.f, 0.f};
    inverse_kinematics(qik, v3(r.tip_init[0], r.tip_init[1], r.tip_init[2]), tq);
#pragma unroll
    // joint control keeps joint_state_target (= the joint state after the reset, kuka.py:165) in the motor
    // targets; the motors stay off (zero impulse limit) until the first action either way
    for (int k = 0; k < 7; k++) { e.q[k] = qik[k]; e.qd[k] = 0.0f; e.mt[k] = io.jc ? qik[k] : 0.0f; e.mi[k] = 0.0f; }
#pragma unroll
    for (int k = 7; k < ND; k++) { e.q[k] = GRIPPER_ABS_LIMIT; e.qd[k] = 0.0f; e.mt[k] = GRIPPER_ABS_LIMIT; e.mi[k] = FINGER_FORCE * OUTER_DT; }
    Frames f;
    forward_kinematics<7>(e.q, f);
    V3 tip = tip_position(f);
    const int spawn_w = 2 * NBLK + io.goal_dim + (io.cur ? 1 : 0);
    float sampled[2 * 5 + 3 * 5 + 4 + 1];
    const float* sp;
    if (r.spawn) sp = r.spawn + (size_t)i * spawn_w;
    else {
      spawn::Philox rng;
      const uint32_t ep = r.episode[i];
      r.episode[i] = ep + 1;
      spawn::stream_init(rng, r.seed, r.env_base + i, ep);
      spawn::sample_row(rng, r.task, NBLK, io.grip_goal, r.bounds, sampled);
      for (int k = 0; k < spawn_w; k++) r.spawn_out[(size_t)i * spawn_w + k] = sampled[k];
      sp = sampled;
    }
#pragma unroll
    for (int b = 0; b < NBLK; b++) {
      e.bpos[b] = v3(sp[2 * b], sp[2 * b + 1], r.spawn_z);
      e.bquat[b][0] = e.bquat[b][1] = e.bquat[b][2] = 0.0f; e.bquat[b][3] = 1.0f;
      e.bv[b] = v3(0, 0, 0); e.bw[b] = v3(0, 0, 0);
    }
#pragma unroll
    for (int k = 0; k < 7; k++) s[(ST_REST + k) * B] = qik[k];
    s[(ST_EE + 0) * B] = tip.x; s[(ST_EE + 1) * B] = tip.y; s[(ST_EE + 2) * B] = tip.z;
    for (int k = 0; k < io.goal_dim; k++) s[(size_t)(ST_BLK + 13 * NBLK + k) * B] = sp[2 * NBLK + k];
    // task decomposition selects the last sub-goal again (kuka_multi_step_base_env.py:247-248); a curriculum reset
    // carries the sub-goal index of the level it drew behind the goal of its spawn row
    if (io.td) s[(size_t)(ST_BLK + 13 * NBLK + io.goal_dim) * B] = io.cur ? sp[2 * NBLK + io.goal_dim] : -1.0f;
    s[(size_t)(io.state_words - 1) * B] = 0.0f;
    store_env<TASK, NBLK>(e, io, i);
  }
  if (r.auto_rows) write_obs<TASK, NBLK, true>(e, io, i);
  else write_obs<TASK, NBLK>(e, io, i);
}

template <int TASK, int NBLK>
__global__ void __launch_bounds__(32) reset_kernel(ResetIO r) {
  using D = Dims<TASK, NBLK>;
  const StepIO& io = r.io;
  const int i = env_of_thread(io);
  if (r.auto_rows) {  // auto-reset pass: only the environments that just finished an episode do anything ...
    const bool mine = i >= 0 && r.mask[i] != 0;
    if (mine) reset_env<TASK, NBLK>(r, i, true);
    if (io.g_n) {  // ... and, when the batch is sharded, the warp then pushes its (final) rows to the peers
      __syncwarp();
      const int lane = threadIdx.x & 31, W = io.row_width;
      const int env0 = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * io.epw;
      const int nenv = max(0, min(io.epw, io.batch - env0));
      const float* rows = io.obs + (size_t)env0 * W;
      for (int d = 0; d < io.g_n; d++) {
        float* rout = io.g_obs[d] + (size_t)env0 * W;
        for (int k = lane; k < nenv * W; k += 32) rout[k] = rows[k];
        if (lane < nenv) { io.g_reward[d][env0 + lane] = io.reward[env0 + lane]; io.g_done[d][env0 + lane] = io.done[env0 + lane]; io.g_success[d][env0 + lane] = io.success[env0 + lane]; }
      }
      __threadfence_system();  // every lane stored (idle ones included): order the stores before any arrival of the warp
      __syncwarp();
      if (i >= 0) gather_arrive(io);
    }
    return;
  }
  if (i < 0) { stage_row(nullptr, io, false); return; }
  reset_env<TASK, NBLK>(r, i, r.mask == nullptr || r.mask[i] != 0);
}


// test aid (pmg_debug_box_box): one box pair per thread through the narrowphase of the step kernels
__global__ void debug_box_box_kernel(const float* in, int64_t n, int stat, float* out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* r = in + 30 * i;
  M3 R1, R2;
  R1.r0 = v3(r[3], r[4], r[5]); R1.r1 = v3(r[6], r[7], r[8]); R1.r2 = v3(r[9], r[10], r[11]);
  R2.r0 = v3(r[18], r[19], r[20]); R2.r1 = v3(r[21], r[22], r[23]); R2.r2 = v3(r[24], r[25], r[26]);
  BoxScratch scr;
  const int nc = box_box(v3(r[0], r[1], r[2]), R1, v3(r[12], r[13], r[14]), v3(r[15], r[16], r[17]), R2, v3(r[27], r[28], r[29]), scr, stat);
  float* o = out + 32 * i;
  o[0] = (float)nc;
  for (int k = 0; k < nc; k++) {
    o[1 + 7 * k] = scr.out[k].pB.x; o[2 + 7 * k] = scr.out[k].pB.y; o[3 + 7 * k] = scr.out[k].pB.z;
    o[4 + 7 * k] = scr.out[k].nB.x; o[5 + 7 * k] = scr.out[k].nB.y; o[6 + 7 * k] = scr.out[k].nB.z; o[7 + 7 * k] = scr.out[k].dist;
  }
}

// packed rows [B, W] -> four contiguous blocks [B, O] [B, P] [B, G] [B, G] (pmg_step_host_blocks)
__global__ void split_rows_kernel(const float* packed, int B, int W, int O, int P, int G, float* blocks) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * W) return;
  const int env = i / W, c = i - env * W;
  size_t dst;
  if (c < O) dst = (size_t)env * O + c;
  else if (c < O + P) dst = (size_t)B * O + (size_t)env * P + (c - O);
  else if (c < O + P + G) dst = (size_t)B * (O + P) + (size_t)env * G + (c - O - P);
  else dst = (size_t)B * (O + P + G) + (size_t)env * G + (c - O - P - G);
  blocks[dst] = packed[i];
}

__global__ void reward_kernel(const float* ag, const float* dg, int64_t n, int g, float thr, int binary, float* reward, uint8_t* ok) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float d2 = 0.0f;
  for (int k = 0; k < g; k++) { float d = ag[i * g + k] - dg[i * g + k]; d2 += d * d; }
  float d = sqrtf(d2);
  bool na = d > thr;
  reward[i] = binary ? -(na ? 1.0f : 0.0f) : -d;
  ok[i] = na ? 0 : 1;
}

// The same computation through shared memory, for goal widths whose rows are a power-of-two number of bytes.
// Measured on the B200 (tools/reward_bw.py, profiles/r01_reward_kernel_bandwidth.txt; algorithmic bytes 8g + 5 per row,
// inputs of 2-3 GB): the row-per-thread kernel above streams at 94 % (g = 3) and 101 % (g = 12) of the measured copy
// bandwidth -- its g strided loads hit the lines the first one brought into L1 -- but only 36 % at g = 16 (64-byte
// rows: the warp's lines fall into a few L1 sets and are evicted before they are re-used).  Here a block walks tiles of
// 256 x RPT rows: the tile is one contiguous run of floats, read with 16-byte loads by all threads, the differences go
// to shared memory (row stride g|1 words: odd, so the per-row sums below are bank-conflict free), then thread r
// accumulates row r in the same order as reward_kernel (bit-identical results): 76 % at g = 16, 69-80 % at g = 3 / 12.
// Used when g is a multiple of 16 (<= 32) and both arrays are 16-byte aligned.
constexpr int REWARD_THREADS = 256;
template <int RPT>  // rows per thread: 256 * RPT rows per tile (small g needs more bytes in flight per block)
__global__ void __launch_bounds__(REWARD_THREADS) reward_kernel_tiled(const float* __restrict__ ag, const float* __restrict__ dg, int64_t n, int g,
                                                                      float thr, int binary, float* __restrict__ reward, uint8_t* __restrict__ ok) {
  constexpr int TILE = REWARD_THREADS * RPT;
  extern __shared__ float diff[];  // TILE * (g | 1)
  const int gp = g | 1;
  const int64_t ntiles = (n + TILE - 1) / TILE;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t row0 = tile * TILE;
    const int rows = (int)(n - row0 < TILE ? n - row0 : TILE);
    const int cnt = rows * g;
    const float* a = ag + row0 * g;  // row0 * g is a multiple of 256 floats: as aligned as the arrays
    const float* b = dg + row0 * g;
    const float4* a4 = reinterpret_cast<const float4*>(a);
    const float4* b4 = reinterpret_cast<const float4*>(b);
    for (int v = threadIdx.x; v < (cnt >> 2); v += REWARD_THREADS) {
      const float4 x = __ldcs(a4 + v), y = __ldcs(b4 + v);
      int r = (4 * v) / g, k = 4 * v - r * g;
      const float d[4] = {x.x - y.x, x.y - y.y, x.z - y.z, x.w - y.w};
#pragma unroll
      for (int j = 0; j < 4; j++) {
        diff[r * gp + k] = d[j];
        if (++k == g) { k = 0; r++; }
      }
    }
    for (int e = (cnt & ~3) + threadIdx.x; e < cnt; e += REWARD_THREADS) {  // the last tile may end inside a 16-byte group
      const int r = e / g;
      diff[r * gp + (e - r * g)] = a[e] - b[e];
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < RPT; j++) {
      const int r = threadIdx.x + j * REWARD_THREADS;
      if (r < rows) {
        const float* dr = diff + r * gp;
        float d2 = 0.0f;
        for (int k = 0; k < g; k++) { const float d = dr[k]; d2 += d * d; }
        const float d = sqrtf(d2);
        const bool na = d > thr;
        reward[row0 + r] = binary ? -(na ? 1.0f : 0.0f) : -d;
        ok[row0 + r] = na ? 0 : 1;
      }
    }
    __syncthreads();
  }
}
template <int RPT>
void launch_reward_tiled(const float* ag, const float* dg, int64_t n, int g, float thr, int binary, float* reward, uint8_t* ok, cudaStream_t st) {
  constexpr int TILE = REWARD_THREADS * RPT;
  const int64_t ntiles = (n + TILE - 1) / TILE;
  const unsigned grid = (unsigned)(ntiles < 148 * 8 ? ntiles : 148 * 8);  // 8 resident blocks of 256 threads per SM
  reward_kernel_tiled<RPT><<<grid, REWARD_THREADS, sizeof(float) * TILE * (g | 1), st>>>(ag, dg, n, g, thr, binary, reward, ok);
}

// ---- hindsight relabelling (see include/pmg.h) ---------------------------------------------------
__host__ __device__ inline uint64_t her_mix(uint64_t z) {  // splitmix64 finaliser
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}
__global__ void her_sample_kernel(int64_t n, int n_episodes, int horizon, float her_prob, uint64_t seed,
                                  int32_t* ep, int32_t* tt, int32_t* fut) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t base = seed + 0x9e3779b97f4a7c15ull * (uint64_t)(3 * i + 1);
  const uint64_t h0 = her_mix(base), h1 = her_mix(base + 0x9e3779b97f4a7c15ull), h2 = her_mix(base + 2 * 0x9e3779b97f4a7c15ull);
  const int e = (int)((h0 >> 32) * (uint64_t)n_episodes >> 32);           // floor(u * n) on the top 32 bits
  const int t = (int)((h1 >> 32) * (uint64_t)horizon >> 32);
  const float u = (float)(h2 >> 40) * (1.0f / 16777216.0f);              // 24-bit uniform in [0, 1)
  const int span = horizon - t;                                           // future indices t + 1 .. horizon
  const int f = u < her_prob ? t + 1 + (int)((h2 & 0xffffffffull) * (uint64_t)span >> 32) : -1;
  ep[i] = e; tt[i] = t; fut[i] = f;
}
__global__ void her_relabel_kernel(const float* ag, const float* dg, int horizon, int g, const int32_t* ep, const int32_t* tt,
                                   const int32_t* fut, int64_t n, float thr, int binary, float* goal_out, float* reward, uint8_t* ok) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int e = ep[i], t = tt[i], f = fut[i];
  const float* achieved = ag + ((size_t)e * (horizon + 1) + t + 1) * g;
  const float* goal = f >= 0 ? ag + ((size_t)e * (horizon + 1) + f) * g : dg + (size_t)e * g;
  float d2 = 0.0f;
  for (int k = 0; k < g; k++) { const float gk = goal[k]; goal_out[i * g + k] = gk; const float d = achieved[k] - gk; d2 += d * d; }
  const float d = sqrtf(d2);
  const bool na = d > thr;
  reward[i] = binary ? -(na ? 1.0f : 0.0f) : -d;
  ok[i] = na ? 0 : 1;
}

// Construction-time state: BaseBulletMGEnv.__init__ resets the robot once on its own (base_env.py:41)
// before its first self.reset(), i.e. the rest pose (kuka.py:27) gets one IK refinement here.
__global__ void init_state_kernel(float* state_base, int batch, int tile, int words, int nblk, float tx, float ty, float tz) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= batch) return;
  const size_t B = tile;  // [tile][word][env in tile], see StepIO::tile
  float* state = state_base + (size_t)(i / tile) * words * tile + i % tile;
  float q[ND];
  for (int k = 0; k < 7; k++) q[k] = c_rest_pose0[k];
  q[7] = q[8] = 0.0f;
  const float tq[4] = {0.f, -1.f, 0.f, 0.f};
  inverse_kinematics(q, v3(tx, ty, tz), tq);
  for (int k = 0; k < 7; k++) { state[(ST_REST + k) * B] = q[k]; state[(ST_Q + k) * B] = q[k]; }
  for (int k = 7; k < ND; k++) state[(ST_Q + k) * B] = GRIPPER_ABS_LIMIT;
  for (int b = 0; b < nblk; b++) state[(ST_BLK + 13 * b + 6) * B] = 1.0f;  // identity quaternion
}

// ------------------------------------------------------------------------------------------------
// host: numpy legacy RandomState (MT19937) -- the reference draws object / goal poses from
// gym's np_random (base_env.py:120-122); same stream => same resets.
// ------------------------------------------------------------------------------------------------
namespace {

struct MT {
  uint32_t mt[624]; int idx;
  void init_genrand(uint32_t s) {
    mt[0] = s;
    for (int i = 1; i < 624; i++) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
    idx = 624;
  }
  void init_by_array(const uint32_t* key, int len) {
    init_genrand(19650218u);
    int i = 1, j = 0, k = 624 > len ? 624 : len;
    for (; k; k--) {
      mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
      if (++i >= 624) { mt[0] = mt[623]; i = 1; }
      if (++j >= len) j = 0;
    }
    for (k = 623; k; k--) {
      mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
      if (++i >= 624) { mt[0] = mt[623]; i = 1; }
    }
    mt[0] = 0x80000000u; idx = 624;
  }
  uint32_t next() {
    if (idx >= 624) {
      for (int k = 0; k < 624; k++) {
        uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
        mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
      }
      idx = 0;
    }
    uint32_t y = mt[idx++];
    y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
    return y;
  }
  double random_sample() { uint32_t a = next() >> 5, b = next() >> 6; return (a * 67108864.0 + b) / 9007199254740992.0; }
  double uniform(double lo, double hi) { return lo + (hi - lo) * random_sample(); }
  uint32_t interval(uint32_t max) {
    if (max == 0) return 0;
    uint32_t mask = max, v;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
    while ((v = next() & mask) > max) {}
    return v;
  }
};

thread_local char g_err[512] = "";
int fail(int code, const char* fmt, const char* detail = "") {
  snprintf(g_err, sizeof g_err, fmt, detail);
  return code;
}
#define CUDA_TRY(expr) do { cudaError_t err__ = (expr); if (err__ != cudaSuccess) return fail(PMG_ERR_CUDA, #expr ": %s", cudaGetErrorString(err__)); } while (0)

}  // namespace

struct pmg_handle {
  pmg_config cfg;
  int nblk, O, P, G, W, A, state_words, man_words, spawn_w;
  int tile = 0;           // environments per tile of the state / manifold arrays (StepIO::tile)
  size_t batch_pad = 0;   // batch rounded up to whole tiles: what the two arrays are allocated for
  size_t state_index(size_t word, size_t env) const { return (env / tile * state_words + word) * tile + env % tile; }
  bool multi = false, grasp = false, grip = false, jc = false, td = false, cur = false;  // task variants, see pmg_config
  // curriculum state, one reference env's worth per environment (kuka_multi_step_base_env.py:122-140)
  bool cur_update = false;
  double cur_goals_per = 0;
  std::vector<double> cur_prob, cur_count;  // [batch][nblk]
  std::vector<int32_t> cur_level;           // [batch]
  float* d_state = nullptr; float* d_man = nullptr; float* d_spawn = nullptr; uint8_t* d_mask = nullptr; int* d_overflow = nullptr; float* d_row_spill = nullptr;
  float* d_blocks = nullptr;
  float* d_action = nullptr; float* d_obs = nullptr; float* d_reward = nullptr; uint8_t* d_done = nullptr; uint8_t* d_success = nullptr;
  float* h_spawn = nullptr;  // pinned; the spawn row every env was last reset with
  float* h_stage[2] = {nullptr, nullptr};  // pinned DMA staging, alternating, so that sampling overlaps the GPU
  cudaEvent_t stage_done[2] = {nullptr, nullptr};
  int stage_cur = 0;
  std::vector<MT> rng;
  double tip_init[3], obj_lo[3], obj_hi[3], tgt_lo[3], tgt_hi[3];
  double spawn_z = 0.175;  // kuka_single_step_base_env.py:50 (cube) / :56 (Slide's puck: 0.170)
  bool was_reset = false;
  int64_t launches = 0;
  int epw = 32;  // environments per warp (launch geometry, see pmg_create)
  bool no_bulk = true;   // TMA staging of the state tile is opt-in (PMG_BULK_COPY=1): measured slower, see DESIGN.md
  bool default_carveout = false;  // PMG_DEFAULT_CARVEOUT=1 keeps the driver's shared-memory carve-out
  bool coop = true;  // Reach: lane-cooperative kernel (PMG_COOP=0 selects the thread-per-env kernel)
  bool coop_block = true;  // Push / PickAndPlace: lane-cooperative kernel (PMG_COOP_BLOCK=0 selects the thread-per-env kernel)
  bool coop_stack = true;  // BlockStack / BlockRearrange with >= 2 blocks: lane-cooperative kernel (PMG_COOP_STACK=0 selects the thread-per-env kernel)
  bool hinted = false;  // shared-memory carve-out hint of this handle's step kernel has been set on its device
  int epb = 0, wpb = 1;  // lane-cooperative kernels: environments per warp and warps per block, chosen at the first launch (coop_geometry)
  // device-side reset sampling (pmg_spawn.cuh) and auto-reset
  bool dev_rng = false, auto_reset = false, last_spawn_on_device = false;
  uint64_t rng_seed = 0; int64_t env_base = 0;
  uint32_t* d_episode = nullptr; float* d_spawn_dev = nullptr; float* terminal_obs = nullptr;
  // gather over peer memory (pmg_gather_*): this rank's buffer, the peers' mapped buffers, layout
  int g_world = 0, g_rank = 0;
  char* g_buf[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // [rank] = own allocation
  bool g_connected = false;
  unsigned g_seq = 0;
  unsigned* d_g_counter = nullptr;
  int* g_err_host = nullptr; int* g_err_dev = nullptr;
  int* bad_host = nullptr; int* bad_dev = nullptr;  // mapped host word the step kernels raise on an out-of-range action
  size_t g_parity_bytes = 0, g_off_reward = 0, g_off_done = 0, g_off_success = 0, g_flags_off = 0, g_total = 0;
  // optional CUDA-event timing of the step kernel alone (pmg_kernel_timing): a ring of event pairs on the launch stream
  bool timing = false;
  std::vector<cudaEvent_t> t_ev;  // 2 * TIMING_RING events, created on first use
  int64_t t_count = 0;
};
constexpr int TIMING_RING = 2048;

namespace {

// level = np_random.choice(num_curriculum, p=curriculum_prob) (kuka_multi_step_envs.py:127,197), which is
// cdf.searchsorted(random_sample(), side='right') in numpy's legacy RandomState
int draw_curriculum_level(pmg_handle* h, int i, MT& r) {
  const int nb = h->nblk;
  const double* prob = &h->cur_prob[(size_t)i * nb];
  double cdf[5], acc = 0;
  for (int k = 0; k < nb; k++) { acc += prob[k]; cdf[k] = acc; }
  const double u = r.random_sample();
  int level = 0;
  while (level < nb && cdf[level] / acc <= u) level++;
  h->cur_level[i] = level;
  return level;
}

// _update_curriculum_prob (kuka_multi_step_base_env.py:350-379), statement by statement, on environment i's own schedule
void update_curriculum_prob(pmg_handle* h, int i, int level) {
  if (!h->cur_update) return;
  const int nb = h->nblk;
  double* prob = &h->cur_prob[(size_t)i * nb];
  double* count = &h->cur_count[(size_t)i * nb];
  count[level] += 1;
  bool fin[5], half[5];
  for (int k = 0; k < nb; k++) {
    fin[k] = count[k] >= h->cur_goals_per; half[k] = count[k] >= h->cur_goals_per / 2;
    if (fin[k]) prob[k] = 0.0;
  }
  if (half[0] && !fin[0]) { prob[0] = 0.5; prob[1] = 0.5; }
  for (int k = 1; k < nb - 1; k++)
    if (fin[k - 1] && !fin[k]) {
      if (half[k]) { prob[k] = 0.5; prob[k + 1] = 0.5; } else prob[k] = 1.0;
    }
  if (fin[nb - 2]) prob[nb - 1] = 1.0;
}

// Sampling of one reset, consuming the env's stream exactly like the reference:
// kuka_single_step_base_env.py:104-148, kuka_multi_step_base_env.py:223-240, kuka_multi_step_envs.py:34-63
void sample_spawn(pmg_handle* h, int i, float* out) {
  MT& r = h->rng[i];
  const int nb = h->nblk;
  double xy[2 * 5];
  if (h->multi) {
    for (int b = 0; b < nb; b++) {
      for (;;) {
        double x = r.uniform(h->obj_lo[0], h->obj_hi[0]), y = r.uniform(h->obj_lo[1], h->obj_hi[1]);
        bool ok = true;
        for (int k = 0; k < b; k++) if (!(hypot(x - xy[2 * k], y - xy[2 * k + 1]) > 0.06)) ok = false;
        if (!(hypot(x - h->tip_init[0], y - h->tip_init[1]) > 0.06)) ok = false;
        if (ok) { xy[2 * b] = x; xy[2 * b + 1] = y; break; }
      }
    }
    for (int b = 0; b < nb; b++) { out[2 * b] = (float)xy[2 * b]; out[2 * b + 1] = (float)xy[2 * b + 1]; }
    if (h->cfg.task == PMG_BLOCK_REARRANGE) {
      // kuka_multi_step_envs.py:174-189: one table target per block, clear of every block and earlier target
      double txy[2 * 5];
      float* goal = out + 2 * nb;
      for (int b = 0; b < nb; b++) {
        for (;;) {
          double x = r.uniform(h->tgt_lo[0], h->tgt_hi[0]), y = r.uniform(h->tgt_lo[1], h->tgt_hi[1]);
          bool ok = true;
          for (int k = 0; k < b; k++) if (!(hypot(x - txy[2 * k], y - txy[2 * k + 1]) > 0.06)) ok = false;
          for (int k = 0; k < nb; k++) if (!(hypot(x - xy[2 * k], y - xy[2 * k + 1]) > 0.06)) ok = false;
          if (ok) { txy[2 * b] = x; txy[2 * b + 1] = y; break; }
        }
        goal[3 * b] = (float)txy[2 * b]; goal[3 * b + 1] = (float)txy[2 * b + 1]; goal[3 * b + 2] = 0.175f;
      }
      if (h->cur) {
        // kuka_multi_step_envs.py:197-225: a level, then the level + 1 blocks to move =
        // sort(np_random.choice(arange(nb), size=level + 1, replace=False)), which numpy's legacy RandomState
        // evaluates as permutation(nb)[:level + 1] (a shuffle of arange(nb)); the moved blocks take the sampled
        // targets in order, the others keep their own position as the goal (rebuilt every observation by the kernels)
        const int level = draw_curriculum_level(h, i, r);
        int perm[5];
        for (int k = 0; k < nb; k++) perm[k] = k;
        for (int k = nb - 1; k > 0; k--) { int j = (int)r.interval((uint32_t)k); int t = perm[k]; perm[k] = perm[j]; perm[j] = t; }
        int moved = 0;
        for (int k = 0; k <= level; k++) moved |= 1 << perm[k];
        update_curriculum_prob(h, i, level);
        int j = 0;
        for (int b = 0; b < nb; b++) {
          const bool mv = (moved >> b) & 1;
          goal[3 * b] = (float)(mv ? txy[2 * j] : xy[2 * b]); goal[3 * b + 1] = (float)(mv ? txy[2 * j + 1] : xy[2 * b + 1]);
          if (mv) j++;
        }
        goal[h->G] = (float)moved;  // consumed by reset_kernel like the stack curriculum's sub-goal index
      }
      return;
    }
    int order[5];
    for (int k = 0; k < nb; k++) order[k] = k;
    for (int k = nb - 1; k > 0; k--) { int j = (int)r.interval((uint32_t)k); int t = order[k]; order[k] = order[j]; order[j] = t; }
    double bx, by;
    for (;;) {
      bx = r.uniform(h->tgt_lo[0], h->tgt_hi[0]); by = r.uniform(h->tgt_lo[1], h->tgt_hi[1]);
      bool ok = true;
      for (int k = 0; k < nb; k++) if (!(hypot(bx - xy[2 * k], by - xy[2 * k + 1]) > 0.08)) ok = false;
      if (ok) break;
    }
    float* goal = out + 2 * nb;
    for (int k = 0; k < nb; k++) {
      goal[3 * order[k]] = (float)bx; goal[3 * order[k] + 1] = (float)by;
      goal[3 * order[k] + 2] = (float)(k == 0 ? 0.175 : 0.175 + 0.03 * k);
    }
    if (h->grip) {  // kuka_multi_step_envs.py:75-77: gripper above the top block, jaws at the grasp width
      goal[3 * nb] = (float)bx; goal[3 * nb + 1] = (float)by;
      goal[3 * nb + 2] = (float)(nb == 1 ? 0.175 : 0.175 + 0.03 * (nb - 1)); goal[3 * nb + 3] = 0.03f;
    }
    if (h->cur) {  // kuka_multi_step_envs.py:127-134
      const int level = draw_curriculum_level(h, i, r);
      goal[h->G] = (float)(h->grip ? 2 * level + 1 : level);  // the equivalent sub-goal index, consumed by reset_kernel
      update_curriculum_prob(h, i, level);
    }
    return;
  }
  double center[3] = {h->tip_init[0], h->tip_init[1], h->tip_init[2]};
  if (nb) {
    double x = h->tip_init[0], y = h->tip_init[1];
    while (hypot(x - h->tip_init[0], y - h->tip_init[1]) < 0.1) {
      x = r.uniform(h->obj_lo[0], h->obj_hi[0]); y = r.uniform(h->obj_lo[1], h->obj_hi[1]);
    }
    out[0] = (float)x; out[1] = (float)y;
    center[0] = x; center[1] = y; center[2] = h->spawn_z;
  }
  double g[3];
  for (;;) {
    for (int k = 0; k < 3; k++) g[k] = r.uniform(h->tgt_lo[k], h->tgt_hi[k]);
    double dx = g[0] - center[0], dy = g[1] - center[1], dz = g[2] - center[2];
    if (sqrt(dx * dx + dy * dy + dz * dz) > 0.1) break;
  }
  if (h->cfg.task == PMG_PUSH || h->cfg.task == PMG_SLIDE) g[2] = h->spawn_z;
  else if (h->cfg.task == PMG_PICK_AND_PLACE) { if (r.uniform(0, 1) >= 0.5) g[2] = 0.175; }
  for (int k = 0; k < 3; k++) out[2 * nb + k] = (float)g[k];
}

StepIO make_io(pmg_handle* h, const float* action, float* obs, float* reward, uint8_t* done, uint8_t* success) {
  StepIO io;
  io.state = h->d_state; io.manifold = h->d_man; io.batch = h->cfg.batch; io.state_words = h->state_words;
  io.tile = h->tile; io.man_words = h->man_words;
  io.action = action; io.obs = obs; io.reward = reward; io.done = done; io.success = success;
  io.thr = h->cfg.distance_threshold; io.binary = h->cfg.binary_reward; io.max_steps = h->cfg.max_episode_steps;
  io.overflow = h->d_overflow; io.row_spill = h->d_row_spill;
  io.epw = h->epw; io.epb = 32 / coop::GL;
  io.bulk = 0; io.tile_offset = 0;
  io.grasp = h->grasp; io.jc = h->jc; io.grip_goal = h->grip; io.td = (h->td || h->cur) ? (h->cfg.task == PMG_BLOCK_REARRANGE ? 2 : 1) : 0; io.cur = h->cur;
  io.adim = h->A; io.goal_dim = h->G; io.row_width = h->W;
  io.g_n = 0; io.g_in_step = 0; io.g_world = 0; io.g_seq = 0; io.g_flags_local = nullptr; io.g_counter = nullptr; io.g_err = nullptr;
  io.bad_action = h->bad_dev;
  return io;
}

// Launch geometry of a lane-cooperative kernel: environments per one-warp block.  A warp of four octets runs as long
// as the union of its octets' paths, and at the configs' batches the kernels are latency-bound (one warp issues every
// 4-5 cycles), so FEWER environments per warp -- more, emptier warps -- is faster as long as (a) every block is still
// resident at once (registers / shared memory: asked from the occupancy API for this kernel and block size) and (b) there
// is at most one warp per scheduler (4 per SM).  Measured (profiles/r02_envs_per_block.txt): pick_and_place at 512
// environments 2.10 -> 1.86 ms with 1 per warp, block_stack at 256: 2.91 -> 2.74 ms; but block_stack at 2048 with 2 per
// warp (7 warps per SM) is SLOWER than 4 per warp (3.71 vs 3.45 ms): more warps in different phases of a loop body larger
// than the instruction cache.  So the shards of the 8-GPU configs (256 - 1024 environments) get emptier warps, the
// single-GPU configs keep 4 per warp.  PMG_COOP_EPB overrides.
template <class K>
int coop_geometry(pmg_handle* h, K kernel, size_t table_bytes, size_t env_bytes, int max_wpb) {
  if (h->epb) return h->epb;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->cfg.device);
  int best = 32 / coop::GL;
  for (int epb = best; epb >= 1; epb >>= 1) {
    const size_t smem = table_bytes + (size_t)epb * env_bytes;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 32, smem) != cudaSuccess) break;
    const long blocks = ((long)h->cfg.batch + epb - 1) / epb;
    if (blocks > (long)per_sm * sms) break;          // (a) one wave
    if (blocks > 4L * sms) break;                    // (b) <= 1 warp per scheduler
    best = epb;
  }
  if (const char* ev = getenv("PMG_COOP_EPB")) { const int v = atoi(ev); if (v == 1 || v == 2 || v == 4) best = v; }
  h->epb = best;
  // Warps per block.  One-warp blocks spread a small batch over all SMs; once an SM holds several warps anyway they
  // go into one block and run the substep loop in lockstep (Grp::block_sync), sharing its instruction stream --
  // as many warps per block as an SM would hold, while the whole batch still fits one wave.
  int wpb = 1, max_optin = 227 * 1024;
  cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->cfg.device);
  const long warps = ((long)h->cfg.batch + best - 1) / best;
  for (int w = max_wpb; w >= 2; w >>= 1) {
    if (warps < (long)w * sms) continue;             // fewer than w warps per SM: leave them on their own SMs
    const size_t smem = table_bytes + (size_t)w * best * env_bytes;
    if (smem > (size_t)max_optin) continue;
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); continue; }
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 32 * w, smem) != cudaSuccess) { cudaGetLastError(); continue; }
    if ((warps + w - 1) / w > (long)per_sm * sms) continue;  // would need a second wave
    wpb = w;
    break;
  }
  if (const char* ev = getenv("PMG_COOP_WPB")) { const int v = atoi(ev); if (v == 1 || ((v == 2 || v == 4) && v <= max_wpb)) wpb = v; }
  if (wpb > 1) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(table_bytes + (size_t)wpb * best * env_bytes));
  h->wpb = wpb;
  return best;
}

// Grid, block and dynamic shared memory of a lane-cooperative launch (after coop_geometry)
struct CoopLaunch { int blocks, threads; size_t smem; };
CoopLaunch coop_launch(const pmg_handle* h, size_t table_bytes, size_t env_bytes) {
  const int per_block = h->epb * h->wpb;
  return {(h->cfg.batch + per_block - 1) / per_block, 32 * h->wpb, table_bytes + (size_t)per_block * env_bytes};
}

// one warp per block: the block scheduler then spreads the (few) warps evenly over the 148 SMs
template <int TASK, int NBLK>
void launch_step(pmg_handle* h, const StepIO& io_in, cudaStream_t st) {
  StepIO io = io_in;
  if (TASK == 0 && h->coop) {
    if (!h->hinted) {  // per handle = per device: function attributes are per device
      if (h->jc) cudaFuncSetAttribute(step_kernel_coop_reach<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
      else cudaFuncSetAttribute(step_kernel_coop_reach<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
      h->hinted = true;
    }
    io.epb = h->jc ? coop_geometry(h, step_kernel_coop_reach<true>, COOP_TABLE_BYTES, coop::env_stride<coop::EnvSmem>(), 2)
                   : coop_geometry(h, step_kernel_coop_reach<false>, COOP_TABLE_BYTES, coop::env_stride<coop::EnvSmem>(), 2);
    const CoopLaunch cl = coop_launch(h, COOP_TABLE_BYTES, coop::env_stride<coop::EnvSmem>());
    if (h->jc) step_kernel_coop_reach<true><<<cl.blocks, cl.threads, cl.smem, st>>>(io);
    else step_kernel_coop_reach<false><<<cl.blocks, cl.threads, cl.smem, st>>>(io);
    return;
  }
  if constexpr (TASK == 3 && NBLK >= 2) {
    if (h->coop_stack && !h->jc) {
      if (!h->hinted) {
        cudaFuncSetAttribute(step_kernel_coop_multi<NBLK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(COOP_TABLE_BYTES + 4 * coop::env_stride<coop::EnvSmemT<NBLK>>()));
        cudaFuncSetAttribute(step_kernel_coop_multi<NBLK>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        h->hinted = true;
      }
      io.epb = coop_geometry(h, step_kernel_coop_multi<NBLK>, COOP_TABLE_BYTES, coop::env_stride<coop::EnvSmemT<NBLK>>(), 4);
      const CoopLaunch cl = coop_launch(h, COOP_TABLE_BYTES, coop::env_stride<coop::EnvSmemT<NBLK>>());
      step_kernel_coop_multi<NBLK><<<cl.blocks, cl.threads, cl.smem, st>>>(io);
      return;
    }
  }
  int warps = (h->cfg.batch + h->epw - 1) / h->epw;
  size_t stage_floats = (size_t)h->epw * h->W;
  io.bulk = (h->epw == 32 && h->cfg.batch % 32 == 0 && !h->no_bulk && !h->grip && !h->td && !h->cur && h->tile != 32 / coop::GL) ? 1 : 0;
  io.tile_offset = (int)((stage_floats + 31) / 32 * 32);
  size_t smem = (io.bulk ? io.tile_offset + (size_t)Dims<TASK, NBLK>::STATE * 32 : stage_floats) * sizeof(float);
  if ((TASK == 1 || TASK == 2) && h->coop_block) {
    if (!h->hinted) {
      if (h->cfg.task == PMG_SLIDE) cudaFuncSetAttribute(step_kernel_coop_block<5>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
      else cudaFuncSetAttribute(step_kernel_coop_block<TASK == 2 ? 2 : 1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
      h->hinted = true;
    }
    if (h->cfg.task == PMG_SLIDE) {  // Push's layout with the long table and the puck (EnvSmemT<1, true>)
      io.epb = coop_geometry(h, step_kernel_coop_block<5>, COOP_TABLE_BYTES, coop::env_stride<coop::EnvSmemT<1, true>>(), 2);
      const CoopLaunch cl = coop_launch(h, COOP_TABLE_BYTES, coop::env_stride<coop::EnvSmemT<1, true>>());
      step_kernel_coop_block<5><<<cl.blocks, cl.threads, cl.smem, st>>>(io);
      return;
    }
    io.epb = coop_geometry(h, step_kernel_coop_block<TASK == 2 ? 2 : 1>, COOP_TABLE_BYTES, coop::env_stride<coop::EnvSmemT<1>>(), 2);
    const CoopLaunch cl = coop_launch(h, COOP_TABLE_BYTES, coop::env_stride<coop::EnvSmemT<1>>());
    step_kernel_coop_block<TASK == 2 ? 2 : 1><<<cl.blocks, cl.threads, cl.smem, st>>>(io);
    return;
  }
  // The per-thread scratch lives in L1-cached local memory: ask for the smallest shared-memory
  // carve-out instead of one sized for the register-limited 8 blocks per SM.
  if (!h->hinted && !h->default_carveout) {  // per handle = per device: function attributes are per device
    cudaFuncSetAttribute(step_kernel<TASK, NBLK>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxL1);
    h->hinted = true;
  }
  step_kernel<TASK, NBLK><<<warps, 32, smem, st>>>(io);
}
template <int TASK, int NBLK>
void launch_reset(pmg_handle* h, const ResetIO& r, cudaStream_t st) {
  int warps = (h->cfg.batch + h->epw - 1) / h->epw;
  size_t smem = (size_t)h->epw * h->W * sizeof(float);
  reset_kernel<TASK, NBLK><<<warps, 32, smem, st>>>(r);
}

#define PMG_DISPATCH(FN, ...)                                                        \
  switch (h->cfg.task) {                                                             \
    case PMG_REACH: FN<0, 0>(__VA_ARGS__); break;                                    \
    case PMG_PUSH: case PMG_SLIDE: FN<1, 1>(__VA_ARGS__); break;                     \
    case PMG_PICK_AND_PLACE: FN<2, 1>(__VA_ARGS__); break;                           \
    default:                                                                         \
      switch (h->nblk) {                                                             \
        case 1: FN<3, 1>(__VA_ARGS__); break;                                        \
        case 2: FN<3, 2>(__VA_ARGS__); break;                                        \
        case 3: FN<3, 3>(__VA_ARGS__); break;                                        \
        case 4: FN<3, 4>(__VA_ARGS__); break;                                        \
        default: FN<3, 5>(__VA_ARGS__); break;                                       \
      }                                                                              \
  }

// Enqueues the reset kernel.  spawn_dev == nullptr: rows sampled on the device.  auto_rows: see ResetIO.
void enqueue_reset(pmg_handle* h, const uint8_t* mask_dev, const float* spawn_dev, float* obs_dev, int auto_rows, cudaStream_t st,
                   const StepIO* step_io = nullptr) {
  ResetIO r;
  r.io = step_io ? *step_io : make_io(h, nullptr, obs_dev, nullptr, nullptr, nullptr);
  r.mask = mask_dev;
  r.spawn = spawn_dev;
  for (int k = 0; k < 3; k++) r.tip_init[k] = (float)h->tip_init[k];
  r.spawn_z = (float)h->spawn_z;
  r.task = h->cfg.task; r.auto_rows = auto_rows; r.seed = h->rng_seed; r.env_base = h->env_base;
  r.episode = h->d_episode; r.spawn_out = h->d_spawn_dev; r.terminal = auto_rows ? h->terminal_obs : nullptr;
  r.bounds = spawn::to_bounds(h->tip_init, h->obj_lo, h->obj_hi, h->tgt_lo, h->tgt_hi);
  PMG_DISPATCH(launch_reset, h, r, st);
  h->launches++;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// C-ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

int pmg_abi_version(void) { return PMG_ABI_VERSION; }
const char* pmg_last_error(void) { return g_err; }

int pmg_create(const pmg_config* cfg, pmg_handle** out) {
  if (!cfg || !out) return fail(PMG_ERR_INVALID, "pmg_create: null argument%s");
  if (cfg->task < PMG_REACH || cfg->task > PMG_SLIDE) return fail(PMG_ERR_INVALID, "pmg_create: invalid task id%s");
  const bool multi = cfg->task == PMG_BLOCK_STACK || cfg->task == PMG_BLOCK_REARRANGE;
  if (multi && (cfg->num_block < 1 || cfg->num_block > 5)) return fail(PMG_ERR_INVALID, "pmg_create: only support up to 5 blocks%s");
  if (cfg->grip_informed_goal && cfg->task != PMG_BLOCK_STACK) return fail(PMG_ERR_INVALID, "pmg_create: grip_informed_goal is a block_stack option%s");
  if (cfg->task_decomposition && cfg->task != PMG_BLOCK_STACK) return fail(PMG_ERR_INVALID, "pmg_create: task_decomposition is a block_stack option%s");
  if (cfg->use_curriculum && (!multi || cfg->num_block < 2)) return fail(PMG_ERR_INVALID, "pmg_create: use_curriculum needs block_stack / block_rearrange with at least 2 blocks%s");
  if (cfg->use_curriculum && cfg->task_decomposition) return fail(PMG_ERR_INVALID, "pmg_create: if using curriculum, task decomposition should be False, vice versa%s");
  if (cfg->batch < 1) return fail(PMG_ERR_INVALID, "pmg_create: batch must be >= 1%s");
  if (cfg->max_episode_steps < 1 || cfg->max_episode_steps >= (1 << 24)) return fail(PMG_ERR_INVALID, "pmg_create: max_episode_steps out of range%s");
  int ndev = 0;
  CUDA_TRY(cudaGetDeviceCount(&ndev));
  if (cfg->device < 0 || cfg->device >= ndev) return fail(PMG_ERR_INVALID, "pmg_create: no such CUDA device%s");
  CUDA_TRY(cudaSetDevice(cfg->device));
  pmg_handle* h = new (std::nothrow) pmg_handle();
  if (!h) return fail(PMG_ERR_INVALID, "pmg_create: out of host memory%s");
  h->cfg = *cfg;
  const int t = cfg->task;
  h->multi = multi;
  h->grasp = t == PMG_PICK_AND_PLACE || t == PMG_BLOCK_STACK;  // kuka_single_step_envs.py:16, kuka_multi_step_envs.py:30,170
  h->grip = cfg->grip_informed_goal != 0;
  h->jc = cfg->joint_control != 0;
  h->td = cfg->task_decomposition != 0;
  h->cur = cfg->use_curriculum != 0;
  h->nblk = t == PMG_REACH ? 0 : (multi ? cfg->num_block : 1);
  h->O = (t == PMG_REACH ? 3 : (multi ? 8 + 16 * h->nblk : 20)) + (h->jc ? 7 : 0);
  h->P = (t == PMG_REACH ? 3 : (multi ? 4 + 3 * h->nblk : 7)) + (h->jc ? 7 : 0);
  h->G = (multi ? 3 * h->nblk : 3) + (h->grip ? 4 : 0);
  h->W = h->O + h->P + 2 * h->G;
  h->A = h->jc ? (h->grasp ? 8 : 7) : (h->grasp ? 4 : 3);  // kuka.py:104-118
  h->state_words = ST_BLK + 13 * h->nblk + h->G + ((h->td || h->cur) ? 1 : 0) + 1;
  h->man_words = num_pairs(h->nblk) * MAN_WORDS;
  h->spawn_w = 2 * h->nblk + h->G + (h->cur ? 1 : 0);
  // kuka.py:35-51 with obj_range = target_range = 0.15 (kuka_single_step_envs.py, kuka_multi_step_envs.py:29)
  spawn::task_bounds(t, h->tip_init, h->obj_lo, h->obj_hi, h->tgt_lo, h->tgt_hi);
  if (t == PMG_SLIDE) h->spawn_z = PMG_PUCK_SPAWN_Z;
  const size_t B = cfg->batch;
  {
    // Launch geometry: lanes [0, epw) of every warp own one environment.  Spreading a batch over
    // more warps (epw < 32) was measured to be slower at every shipped size (it multiplies the L1
    // footprint of the per-thread scratch), so the default is a full warp; PMG_ENVS_PER_WARP
    // overrides it for experiments.
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, cfg->device));
    const long target_warps = 1;  // measured: 32 envs/warp is fastest at every shipped batch (profiles/)
    int epw = 32;
    while (epw > 1 && (long)(B + epw - 1) / epw < target_warps) epw >>= 1;
    if (const char* ev = getenv("PMG_ENVS_PER_WARP")) { int v = atoi(ev); if (v >= 1 && v <= 32) epw = v; }
    h->epw = epw;
    if (const char* ev = getenv("PMG_BULK_COPY")) h->no_bulk = atoi(ev) == 0;
    if (const char* ev = getenv("PMG_DEFAULT_CARVEOUT")) h->default_carveout = atoi(ev) != 0;
    if (const char* ev = getenv("PMG_COOP")) h->coop = atoi(ev) != 0;
    if (const char* ev = getenv("PMG_COOP_BLOCK")) h->coop_block = atoi(ev) != 0;
    if (const char* ev = getenv("PMG_COOP_STACK")) h->coop_stack = atoi(ev) != 0;
  }
  if (h->cur) {
    h->cur_goals_per = (double)(cfg->num_goals_to_generate / h->nblk);  // floor division (kuka_multi_step_base_env.py:138)
    h->cur_prob.assign(B * h->nblk, 0.0);
    h->cur_count.assign(B * h->nblk, 0.0);
    h->cur_level.assign(B, 0);
    for (size_t i = 0; i < B; i++) h->cur_prob[i * h->nblk] = 1.0;  // the easiest goal is the only possible one at first
  }
  h->rng.resize(B);
  for (size_t i = 0; i < B; i++) h->rng[i].init_genrand(5489u + (uint32_t)i);
#define ALLOC(ptr, bytes) do { cudaError_t e_ = cudaMalloc((void**)&(ptr), (bytes)); if (e_ != cudaSuccess) { pmg_destroy(h); return fail(PMG_ERR_CUDA, "cudaMalloc: %s", cudaGetErrorString(e_)); } } while (0)
  {
    // which kernel family steps this handle (the conditions of launch_step): 4-environment tiles for the lane-cooperative
    // kernels, 32 for the thread-per-env ones, PMG_STATE_TILE=0: the plain [word][env] arrays
    const int t = h->cfg.task;
    const bool coop_handle = (t == PMG_REACH && h->coop) || ((t == PMG_PUSH || t == PMG_PICK_AND_PLACE || t == PMG_SLIDE) && h->coop_block) ||
                             (h->multi && h->nblk >= 2 && h->coop_stack && !h->jc);
    h->tile = coop_handle ? 32 / coop::GL : 32;
    if (const char* ev = getenv("PMG_STATE_TILE")) { if (atoi(ev) == 0) h->tile = (int)B; }
    h->batch_pad = (B + h->tile - 1) / h->tile * h->tile;
  }
  ALLOC(h->d_state, sizeof(float) * h->state_words * h->batch_pad);
  ALLOC(h->d_man, sizeof(float) * h->man_words * h->batch_pad);
  ALLOC(h->d_spawn, sizeof(float) * h->spawn_w * B);
  ALLOC(h->d_mask, B);
  ALLOC(h->d_overflow, sizeof(int));
  if (cudaHostAlloc((void**)&h->bad_host, sizeof(int), cudaHostAllocMapped) != cudaSuccess ||
      cudaHostGetDevicePointer((void**)&h->bad_dev, h->bad_host, 0) != cudaSuccess) { pmg_destroy(h); return fail(PMG_ERR_CUDA, "pmg_create: cannot map the action-error word%s"); }
  *h->bad_host = 0;
  ALLOC(h->d_episode, sizeof(uint32_t) * B);
  ALLOC(h->d_spawn_dev, sizeof(float) * h->spawn_w * B);
  if (h->cfg.task == PMG_SLIDE && !h->coop_block) { pmg_destroy(h); return fail(PMG_ERR_INVALID, "pmg_create: slide runs on the lane-cooperative kernel only (PMG_COOP_BLOCK=0 is set)%s"); }
  if ((h->cfg.task == PMG_PUSH || h->cfg.task == PMG_PICK_AND_PLACE || h->cfg.task == PMG_SLIDE) && h->coop_block)
    ALLOC(h->d_row_spill, sizeof(float) * coop::EnvSmemT<1>::SPILL_WORDS * B);
  if (h->multi && h->nblk >= 2 && h->coop_stack && !h->jc)
    ALLOC(h->d_row_spill, sizeof(float) * coop::EnvSmemT<2>::SPILL_WORDS * B);  // the same for every NBLK >= 2
  ALLOC(h->d_action, sizeof(float) * h->A * B);
  ALLOC(h->d_obs, sizeof(float) * h->W * B);
  ALLOC(h->d_blocks, sizeof(float) * h->W * B);
  ALLOC(h->d_reward, sizeof(float) * B);
  ALLOC(h->d_done, B);
  ALLOC(h->d_success, B);
#undef ALLOC
  if (cudaMallocHost((void**)&h->h_spawn, sizeof(float) * h->spawn_w * B) != cudaSuccess) { pmg_destroy(h); return fail(PMG_ERR_CUDA, "cudaMallocHost failed%s"); }
  for (int k = 0; k < 2; k++)
    if (cudaMallocHost((void**)&h->h_stage[k], sizeof(float) * h->spawn_w * B) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->stage_done[k], cudaEventDisableTiming) != cudaSuccess) { pmg_destroy(h); return fail(PMG_ERR_CUDA, "cudaMallocHost failed%s"); }
  memset(h->h_spawn, 0, sizeof(float) * h->spawn_w * B);
  cudaMemset(h->d_state, 0, sizeof(float) * h->state_words * h->batch_pad);
  cudaMemset(h->d_man, 0, sizeof(float) * h->man_words * h->batch_pad);
  cudaMemset(h->d_overflow, 0, sizeof(int));
  cudaMemset(h->d_episode, 0, sizeof(uint32_t) * B);
  cudaMemset(h->d_spawn_dev, 0, sizeof(float) * h->spawn_w * B);
  fill_lane_table_kernel<<<1, 32>>>();  // per device (a __device__ array): cheap enough to redo per handle
  init_state_kernel<<<(int)((B + 31) / 32), 32>>>(h->d_state, (int)B, h->tile, h->state_words, h->nblk, (float)h->tip_init[0], (float)h->tip_init[1], (float)h->tip_init[2]);
  h->launches++;
  CUDA_TRY(cudaDeviceSynchronize());
  *out = h;
  return PMG_OK;
}

int pmg_destroy(pmg_handle* h) {
  if (!h) return PMG_OK;
  cudaSetDevice(h->cfg.device);
  cudaFree(h->d_state); cudaFree(h->d_man); cudaFree(h->d_spawn); cudaFree(h->d_mask); cudaFree(h->d_overflow); cudaFree(h->d_row_spill);
  cudaFree(h->d_episode); cudaFree(h->d_spawn_dev);
  for (int d = 0; d < h->g_world; d++) {
    if (!h->g_buf[d]) continue;
    if (d == h->g_rank) cudaFree(h->g_buf[d]); else cudaIpcCloseMemHandle(h->g_buf[d]);
  }
  cudaFree(h->d_g_counter);
  if (h->g_err_host) cudaFreeHost(h->g_err_host);
  if (h->bad_host) cudaFreeHost(h->bad_host);
  for (auto& e : h->t_ev) cudaEventDestroy(e);
  cudaFree(h->d_action); cudaFree(h->d_obs); cudaFree(h->d_blocks); cudaFree(h->d_reward); cudaFree(h->d_done); cudaFree(h->d_success);
  if (h->h_spawn) cudaFreeHost(h->h_spawn);
  for (int k = 0; k < 2; k++) { if (h->h_stage[k]) cudaFreeHost(h->h_stage[k]); if (h->stage_done[k]) cudaEventDestroy(h->stage_done[k]); }
  delete h;
  return PMG_OK;
}

int pmg_dims(const pmg_handle* h, int32_t dims[6]) {
  if (!h || !dims) return fail(PMG_ERR_INVALID, "pmg_dims: null argument%s");
  dims[0] = h->O; dims[1] = h->P; dims[2] = h->G; dims[3] = h->G; dims[4] = h->A; dims[5] = h->W;
  return PMG_OK;
}

int pmg_seed(pmg_handle* h, const uint32_t* keys, const int32_t* lens, int32_t max_len) {
  if (!h || !keys || !lens || max_len < 1) return fail(PMG_ERR_INVALID, "pmg_seed: bad argument%s");
  for (int i = 0; i < h->cfg.batch; i++) {
    if (lens[i] < 1 || lens[i] > max_len) return fail(PMG_ERR_INVALID, "pmg_seed: key length out of range%s");
    h->rng[i].init_by_array(keys + (size_t)i * max_len, lens[i]);
  }
  return PMG_OK;
}

int pmg_spawn_width(const pmg_handle* h) { return h ? h->spawn_w : PMG_ERR_INVALID; }

int pmg_last_spawn(const pmg_handle* h, float* spawn_host) {
  if (!h || !spawn_host) return fail(PMG_ERR_INVALID, "pmg_last_spawn: null argument%s");
  if (h->last_spawn_on_device) {  // rows sampled by the reset kernel (pmg_reset_device / auto-reset)
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpy(spawn_host, h->d_spawn_dev, sizeof(float) * h->spawn_w * h->cfg.batch, cudaMemcpyDeviceToHost));
    return PMG_OK;
  }
  memcpy(spawn_host, h->h_spawn, sizeof(float) * h->spawn_w * h->cfg.batch);
  return PMG_OK;
}

int pmg_reset(pmg_handle* h, const uint8_t* mask_host, const float* spawn_host, float* obs_dev, void* stream) {
  if (!h || !obs_dev) return fail(PMG_ERR_INVALID, "pmg_reset: null argument%s");
  cudaStream_t st = (cudaStream_t)stream;
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  const size_t B = h->cfg.batch;
  // host sampling runs while the GPU is still busy with the step kernel enqueued before this reset
  for (size_t i = 0; i < B; i++) {
    if (mask_host && !mask_host[i]) continue;
    if (spawn_host) {
      memcpy(h->h_spawn + i * h->spawn_w, spawn_host + i * h->spawn_w, sizeof(float) * h->spawn_w);
      if (h->cur) {
        const int ind = (int)spawn_host[i * h->spawn_w + h->spawn_w - 1];
        h->cur_level[i] = h->cfg.task == PMG_BLOCK_REARRANGE ? __builtin_popcount((unsigned)ind) - 1 : (h->grip ? ind >> 1 : ind);  // rearrange: the mask of moved blocks
      }
    }
    else sample_spawn(h, (int)i, h->h_spawn + i * h->spawn_w);
  }
  // two pinned staging buffers alternate; one is reused only after the copy that last read it has completed
  float* stage = h->h_stage[h->stage_cur];
  CUDA_TRY(cudaEventSynchronize(h->stage_done[h->stage_cur]));
  memcpy(stage, h->h_spawn, sizeof(float) * h->spawn_w * B);
  CUDA_TRY(cudaMemcpyAsync(h->d_spawn, stage, sizeof(float) * h->spawn_w * B, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaEventRecord(h->stage_done[h->stage_cur], st));
  h->stage_cur ^= 1;
  if (mask_host) CUDA_TRY(cudaMemcpyAsync(h->d_mask, mask_host, B, cudaMemcpyHostToDevice, st));
  enqueue_reset(h, mask_host ? h->d_mask : nullptr, h->d_spawn, obs_dev, 0, st);
  CUDA_TRY(cudaGetLastError());
  if (mask_host) CUDA_TRY(cudaStreamSynchronize(st));  // mask_host may be pageable caller memory
  h->was_reset = true;
  h->last_spawn_on_device = false;
  return PMG_OK;
}

int pmg_set_device_rng(pmg_handle* h, uint64_t seed, int64_t env_index_base) {
  if (!h) return fail(PMG_ERR_INVALID, "pmg_set_device_rng: null handle%s");
  if (h->cur) return fail(PMG_ERR_STATE, "pmg_set_device_rng: curriculum resets are sampled on the host (their schedule lives there)%s");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  h->dev_rng = true; h->rng_seed = seed; h->env_base = env_index_base;
  CUDA_TRY(cudaDeviceSynchronize());
  CUDA_TRY(cudaMemset(h->d_episode, 0, sizeof(uint32_t) * h->cfg.batch));
  return PMG_OK;
}

int pmg_reset_device(pmg_handle* h, const uint8_t* mask_dev, float* obs_dev, void* stream) {
  if (!h || !obs_dev) return fail(PMG_ERR_INVALID, "pmg_reset_device: null argument%s");
  if (!h->dev_rng) return fail(PMG_ERR_STATE, "pmg_reset_device: call pmg_set_device_rng first%s");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  enqueue_reset(h, mask_dev, nullptr, obs_dev, 0, (cudaStream_t)stream);
  CUDA_TRY(cudaGetLastError());
  h->was_reset = true;
  h->last_spawn_on_device = true;
  return PMG_OK;
}

int pmg_set_auto_reset(pmg_handle* h, int32_t on, float* terminal_obs_dev) {
  if (!h) return fail(PMG_ERR_INVALID, "pmg_set_auto_reset: null handle%s");
  if (on && !h->dev_rng) return fail(PMG_ERR_STATE, "pmg_set_auto_reset: call pmg_set_device_rng first%s");
  h->auto_reset = on != 0;
  h->terminal_obs = on ? terminal_obs_dev : nullptr;
  return PMG_OK;
}

int pmg_set_curriculum_update(pmg_handle* h, int32_t on) {
  if (!h) return fail(PMG_ERR_INVALID, "pmg_set_curriculum_update: null handle%s");
  if (!h->cur) return fail(PMG_ERR_STATE, "pmg_set_curriculum_update: the handle was created without use_curriculum%s");
  h->cur_update = on != 0;
  return PMG_OK;
}

int pmg_get_curriculum(const pmg_handle* h, float* prob_host, int32_t* level_host) {
  if (!h) return fail(PMG_ERR_INVALID, "pmg_get_curriculum: null handle%s");
  if (!h->cur) return fail(PMG_ERR_STATE, "pmg_get_curriculum: the handle was created without use_curriculum%s");
  if (prob_host) for (size_t k = 0; k < h->cur_prob.size(); k++) prob_host[k] = (float)h->cur_prob[k];
  if (level_host) memcpy(level_host, h->cur_level.data(), sizeof(int32_t) * h->cur_level.size());
  return PMG_OK;
}

int pmg_set_sub_goal(pmg_handle* h, const int32_t* ind_host, void* stream) {
  if (!h) return fail(PMG_ERR_INVALID, "pmg_set_sub_goal: null handle%s");
  if (!h->td) return fail(PMG_ERR_STATE, "pmg_set_sub_goal: the handle was created without task_decomposition%s");
  cudaStream_t st = (cudaStream_t)stream;
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  const size_t B = h->cfg.batch;
  const int nsub = h->grip ? 2 * h->nblk : h->nblk;
  std::vector<float> v(B, -1.0f);
  if (ind_host)
    for (size_t i = 0; i < B; i++) {
      if (ind_host[i] < -nsub || ind_host[i] >= nsub) return fail(PMG_ERR_INVALID, "pmg_set_sub_goal: list index out of range%s");
      v[i] = (float)ind_host[i];
    }
  // one word of every environment: runs of `tile` floats, one per tile, (state_words * tile) floats apart
  const size_t word = (size_t)ST_BLK + 13 * h->nblk + h->G;
  v.resize(h->batch_pad, -1.0f);
  CUDA_TRY(cudaMemcpy2DAsync(h->d_state + word * h->tile, sizeof(float) * h->state_words * h->tile, v.data(), sizeof(float) * h->tile,
                             sizeof(float) * h->tile, h->batch_pad / h->tile, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaStreamSynchronize(st));  // v is a temporary
  return PMG_OK;
}

int pmg_step(pmg_handle* h, const float* action_dev, float* obs_dev, float* reward_dev, uint8_t* done_dev, uint8_t* success_dev, void* stream) {
  if (!h || !action_dev || !obs_dev || !reward_dev || !done_dev || !success_dev) return fail(PMG_ERR_INVALID, "pmg_step: null argument%s");
  if (!h->was_reset) return fail(PMG_ERR_STATE, "pmg_step: call pmg_reset first%s");
  cudaStream_t st = (cudaStream_t)stream;
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  StepIO io = make_io(h, action_dev, obs_dev, reward_dev, done_dev, success_dev);
  if (h->timing) cudaEventRecord(h->t_ev[2 * (h->t_count % TIMING_RING)], st);
  PMG_DISPATCH(launch_step, h, io, st);
  if (h->timing) { cudaEventRecord(h->t_ev[2 * (h->t_count % TIMING_RING) + 1], st); h->t_count++; }
  h->launches++;
  if (h->auto_reset) {  // environments whose episode just ended reset themselves; their rows become the new episode's first observation
    enqueue_reset(h, done_dev, nullptr, obs_dev, 1, st);
    h->last_spawn_on_device = true;
  }
  CUDA_TRY(cudaGetLastError());
  return PMG_OK;
}

// ---- multi-GPU gather over peer memory ----------------------------------------------------------------------------
int pmg_gather_create(pmg_handle* h, int32_t rank, int32_t world, void* ipc_handle_out) {
  if (!h || !ipc_handle_out) return fail(PMG_ERR_INVALID, "pmg_gather_create: null argument%s");
  if (world < 1 || world > 8 || rank < 0 || rank >= world) return fail(PMG_ERR_INVALID, "pmg_gather_create: 1 <= world <= 8, 0 <= rank < world%s");
  if (h->g_world) return fail(PMG_ERR_STATE, "pmg_gather_create: the handle already has a gather buffer%s");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "include/pmg.h documents 64-byte IPC handles");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  const size_t Bg = (size_t)h->cfg.batch * world;
  auto up = [](size_t v) { return (v + 255) / 256 * 256; };
  h->g_off_reward = up(Bg * h->W * sizeof(float));
  h->g_off_done = h->g_off_reward + up(Bg * sizeof(float));
  h->g_off_success = h->g_off_done + up(Bg);
  h->g_parity_bytes = h->g_off_success + up(Bg);
  h->g_flags_off = 2 * h->g_parity_bytes;
  h->g_total = h->g_flags_off + 256;
  char* buf = nullptr;
  CUDA_TRY(cudaMalloc((void**)&buf, h->g_total));
  CUDA_TRY(cudaMemset(buf, 0, h->g_total));
  CUDA_TRY(cudaMalloc((void**)&h->d_g_counter, sizeof(unsigned)));
  CUDA_TRY(cudaMemset(h->d_g_counter, 0, sizeof(unsigned)));
  CUDA_TRY(cudaHostAlloc((void**)&h->g_err_host, sizeof(int), cudaHostAllocMapped));
  *h->g_err_host = 0;
  CUDA_TRY(cudaHostGetDevicePointer((void**)&h->g_err_dev, h->g_err_host, 0));
  CUDA_TRY(cudaDeviceSynchronize());
  h->g_world = world; h->g_rank = rank; h->g_buf[rank] = buf; h->g_seq = 0;
  h->g_connected = world == 1;
  CUDA_TRY(cudaIpcGetMemHandle((cudaIpcMemHandle_t*)ipc_handle_out, buf));
  return PMG_OK;
}

int pmg_gather_connect(pmg_handle* h, const void* ipc_handles) {
  if (!h || !ipc_handles) return fail(PMG_ERR_INVALID, "pmg_gather_connect: null argument%s");
  if (!h->g_world) return fail(PMG_ERR_STATE, "pmg_gather_connect: call pmg_gather_create first%s");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  const cudaIpcMemHandle_t* hs = (const cudaIpcMemHandle_t*)ipc_handles;
  for (int d = 0; d < h->g_world; d++) {
    if (d == h->g_rank || h->g_buf[d]) continue;
    void* p = nullptr;
    CUDA_TRY(cudaIpcOpenMemHandle(&p, hs[d], cudaIpcMemLazyEnablePeerAccess));
    h->g_buf[d] = (char*)p;
  }
  h->g_connected = true;
  return PMG_OK;
}

int pmg_gather_layout(const pmg_handle* h, int64_t layout[6]) {
  if (!h || !layout) return fail(PMG_ERR_INVALID, "pmg_gather_layout: null argument%s");
  if (!h->g_world) return fail(PMG_ERR_STATE, "pmg_gather_layout: call pmg_gather_create first%s");
  layout[0] = (int64_t)h->g_parity_bytes; layout[1] = 0; layout[2] = (int64_t)h->g_off_reward; layout[3] = (int64_t)h->g_off_done;
  layout[4] = (int64_t)h->g_off_success; layout[5] = (int64_t)h->g_total;
  return PMG_OK;
}

int pmg_step_gather(pmg_handle* h, const float* action_dev, void** gathered_dev_out, void* stream) {
  if (!h || !action_dev || !gathered_dev_out) return fail(PMG_ERR_INVALID, "pmg_step_gather: null argument%s");
  if (!h->was_reset) return fail(PMG_ERR_STATE, "pmg_step_gather: call pmg_reset first%s");
  if (!h->g_connected) return fail(PMG_ERR_STATE, "pmg_step_gather: call pmg_gather_create and pmg_gather_connect first%s");
  if (*h->g_err_host) return fail(PMG_ERR_CUDA, "pmg_step_gather: a peer rank did not publish its step within 5 s%s");
  cudaStream_t st = (cudaStream_t)stream;
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  const unsigned seq = ++h->g_seq;
  const size_t par = (size_t)(seq & 1u) * h->g_parity_bytes;
  const size_t B = h->cfg.batch, row0 = B * h->g_rank;
  auto slice = [&](int d, float*& obs, float*& rew, uint8_t*& dn, uint8_t*& su) {
    char* base = h->g_buf[d] + par;
    obs = (float*)base + row0 * h->W; rew = (float*)(base + h->g_off_reward) + row0;
    dn = (uint8_t*)(base + h->g_off_done) + row0; su = (uint8_t*)(base + h->g_off_success) + row0;
  };
  float *obs, *rew; uint8_t *dn, *su;
  slice(h->g_rank, obs, rew, dn, su);
  StepIO io = make_io(h, action_dev, obs, rew, dn, su);
  int n = 0;
  for (int d = 0; d < h->g_world; d++) {
    io.g_flag[d] = (unsigned*)(h->g_buf[d] + h->g_flags_off) + h->g_rank;
    if (d == h->g_rank) continue;
    slice(d, io.g_obs[n], io.g_reward[n], io.g_done[n], io.g_success[n]);
    n++;
  }
  io.g_n = n; io.g_world = h->g_world; io.g_seq = seq; io.g_in_step = h->auto_reset ? 0 : 1;
  io.g_flags_local = (const unsigned*)(h->g_buf[h->g_rank] + h->g_flags_off);
  io.g_counter = h->d_g_counter; io.g_err = h->g_err_dev;
  if (n == 0) io.g_in_step = 0;  // world 1: nothing to push, nothing to wait for
  if (h->timing) cudaEventRecord(h->t_ev[2 * (h->t_count % TIMING_RING)], st);
  PMG_DISPATCH(launch_step, h, io, st);
  if (h->timing) { cudaEventRecord(h->t_ev[2 * (h->t_count % TIMING_RING) + 1], st); h->t_count++; }
  h->launches++;
  if (h->auto_reset) {  // the reset pass rewrites the rows of the finished environments, then pushes and publishes
    enqueue_reset(h, dn, nullptr, obs, 1, st, &io);
    h->last_spawn_on_device = true;
  }
  CUDA_TRY(cudaGetLastError());
  char* base = h->g_buf[h->g_rank] + par;
  gathered_dev_out[0] = base; gathered_dev_out[1] = base + h->g_off_reward; gathered_dev_out[2] = base + h->g_off_done; gathered_dev_out[3] = base + h->g_off_success;
  return PMG_OK;
}

static bool host_actions_in_box(const float* a, size_t n) {  // kuka.py:168: action_space.contains(a); NaN fails
  for (size_t i = 0; i < n; i++) if (!(a[i] >= -1.0f && a[i] <= 1.0f)) return false;
  return true;
}

int pmg_step_host(pmg_handle* h, const float* action_host, float* obs_host, float* reward_host, uint8_t* done_host, uint8_t* success_host, void* stream) {
  if (!h || !action_host || !obs_host || !reward_host || !done_host || !success_host) return fail(PMG_ERR_INVALID, "pmg_step_host: null argument%s");
  cudaStream_t st = (cudaStream_t)stream;
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  const size_t B = h->cfg.batch;
  if (!host_actions_in_box(action_host, (size_t)h->A * B)) return fail(PMG_ERR_INVALID, "pmg_step_host: action outside the action space Box(-1, 1)%s");
  CUDA_TRY(cudaMemcpyAsync(h->d_action, action_host, sizeof(float) * h->A * B, cudaMemcpyHostToDevice, st));
  int rc = pmg_step(h, h->d_action, h->d_obs, h->d_reward, h->d_done, h->d_success, stream);
  if (rc != PMG_OK) return rc;
  CUDA_TRY(cudaMemcpyAsync(obs_host, h->d_obs, sizeof(float) * h->W * B, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(reward_host, h->d_reward, sizeof(float) * B, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(done_host, h->d_done, B, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(success_host, h->d_success, B, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return PMG_OK;
}

int pmg_step_host_blocks(pmg_handle* h, const float* action_host, float* blocks_host, float* reward_host, uint8_t* done_host,
                         uint8_t* success_host, void* stream) {
  if (!h || !action_host || !blocks_host || !reward_host || !done_host || !success_host) return fail(PMG_ERR_INVALID, "pmg_step_host_blocks: null argument%s");
  cudaStream_t st = (cudaStream_t)stream;
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  const size_t B = h->cfg.batch;
  if (!host_actions_in_box(action_host, (size_t)h->A * B)) return fail(PMG_ERR_INVALID, "pmg_step_host_blocks: action outside the action space Box(-1, 1)%s");
  CUDA_TRY(cudaMemcpyAsync(h->d_action, action_host, sizeof(float) * h->A * B, cudaMemcpyHostToDevice, st));
  int rc = pmg_step(h, h->d_action, h->d_obs, h->d_reward, h->d_done, h->d_success, stream);
  if (rc != PMG_OK) return rc;
  const int n = (int)(B * h->W);
  split_rows_kernel<<<(n + 255) / 256, 256, 0, st>>>(h->d_obs, (int)B, h->W, h->O, h->P, h->G, h->d_blocks);
  h->launches++;
  CUDA_TRY(cudaMemcpyAsync(blocks_host, h->d_blocks, sizeof(float) * h->W * B, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(reward_host, h->d_reward, sizeof(float) * B, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(done_host, h->d_done, B, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(success_host, h->d_success, B, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return PMG_OK;
}

int pmg_compute_reward(const float* ag, const float* dg, int64_t n, int32_t g, float thr, int32_t binary, float* reward, uint8_t* ok, void* stream) {
  if (!ag || !dg || !reward || !ok || n < 0 || g < 1) return fail(PMG_ERR_INVALID, "pmg_compute_reward: bad argument%s");
  if (n == 0) return PMG_OK;
  // A/B runs: PMG_REWARD_SIMPLE=1 / PMG_REWARD_TILED=1 force one kernel wherever it applies
  static const bool simple_only = [] { const char* ev = getenv("PMG_REWARD_SIMPLE"); return ev && atoi(ev) != 0; }();
  static const bool tiled_always = [] { const char* ev = getenv("PMG_REWARD_TILED"); return ev && atoi(ev) != 0; }();
  const bool aligned = (((uintptr_t)ag | (uintptr_t)dg) & 15u) == 0;
  if (aligned && g <= 32 && !simple_only && (g % 16 == 0 || tiled_always)) {
    // <= 27 KB of shared memory per block either way (8 blocks per SM)
    if (g <= 4) launch_reward_tiled<4>(ag, dg, n, g, thr, binary, reward, ok, (cudaStream_t)stream);
    else if (g <= 12) launch_reward_tiled<2>(ag, dg, n, g, thr, binary, reward, ok, (cudaStream_t)stream);
    else launch_reward_tiled<1>(ag, dg, n, g, thr, binary, reward, ok, (cudaStream_t)stream);
  } else {
    reward_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(ag, dg, n, g, thr, binary, reward, ok);
  }
  CUDA_TRY(cudaGetLastError());
  return PMG_OK;
}

int pmg_her_sample(int64_t n, int32_t n_episodes, int32_t horizon, float her_prob, uint64_t seed, int32_t* ep, int32_t* tt,
                   int32_t* fut, void* stream) {
  if (!ep || !tt || !fut || n < 0 || n_episodes < 1 || horizon < 1) return fail(PMG_ERR_INVALID, "pmg_her_sample: bad argument%s");
  if (n == 0) return PMG_OK;
  her_sample_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, n_episodes, horizon, her_prob, seed, ep, tt, fut);
  CUDA_TRY(cudaGetLastError());
  return PMG_OK;
}

int pmg_her_relabel(const float* ag, const float* dg, int32_t n_episodes, int32_t horizon, int32_t g, const int32_t* ep,
                    const int32_t* tt, const int32_t* fut, int64_t n, float thr, int32_t binary, float* goal_out, float* reward,
                    uint8_t* ok, void* stream) {
  if (!ag || !dg || !ep || !tt || !fut || !goal_out || !reward || !ok || n < 0 || n_episodes < 1 || horizon < 1 || g < 1)
    return fail(PMG_ERR_INVALID, "pmg_her_relabel: bad argument%s");
  if (n == 0) return PMG_OK;
  her_relabel_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(ag, dg, horizon, g, ep, tt, fut, n, thr, binary, goal_out, reward, ok);
  CUDA_TRY(cudaGetLastError());
  return PMG_OK;
}

int pmg_state_width(const pmg_handle* h) { return h ? h->state_words : PMG_ERR_INVALID; }

int pmg_get_state(pmg_handle* h, float* out) {
  if (!h || !out) return fail(PMG_ERR_INVALID, "pmg_get_state: null argument%s");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  const size_t B = h->cfg.batch, Wd = h->state_words;
  std::vector<float> tmp(h->batch_pad * Wd);
  CUDA_TRY(cudaDeviceSynchronize());
  CUDA_TRY(cudaMemcpy(tmp.data(), h->d_state, sizeof(float) * h->batch_pad * Wd, cudaMemcpyDeviceToHost));
  for (size_t i = 0; i < B; i++) for (size_t w = 0; w < Wd; w++) out[i * Wd + w] = tmp[h->state_index(w, i)];
  return PMG_OK;
}

int pmg_set_state(pmg_handle* h, const float* in) {
  if (!h || !in) return fail(PMG_ERR_INVALID, "pmg_set_state: null argument%s");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  const size_t B = h->cfg.batch, Wd = h->state_words;
  std::vector<float> tmp(h->batch_pad * Wd, 0.0f);
  for (size_t i = 0; i < B; i++) for (size_t w = 0; w < Wd; w++) tmp[h->state_index(w, i)] = in[i * Wd + w];
  CUDA_TRY(cudaDeviceSynchronize());
  CUDA_TRY(cudaMemcpy(h->d_state, tmp.data(), sizeof(float) * h->batch_pad * Wd, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemset(h->d_man, 0, sizeof(float) * h->man_words * h->batch_pad));
  h->was_reset = true;
  return PMG_OK;
}

int64_t pmg_launch_count(const pmg_handle* h) { return h ? h->launches : 0; }

int pmg_kernel_timing(pmg_handle* h, int32_t on) {
  if (!h) return fail(PMG_ERR_INVALID, "pmg_kernel_timing: null handle%s");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  if (on && h->t_ev.empty()) {
    h->t_ev.resize(2 * TIMING_RING);
    for (auto& e : h->t_ev) CUDA_TRY(cudaEventCreate(&e));
  }
  h->timing = on != 0;
  h->t_count = 0;
  return PMG_OK;
}

int pmg_kernel_time_ms(pmg_handle* h, double* total_ms, int64_t* count) {
  if (!h || !total_ms || !count) return fail(PMG_ERR_INVALID, "pmg_kernel_time_ms: null argument%s");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  CUDA_TRY(cudaDeviceSynchronize());
  const int64_t n = h->t_count < TIMING_RING ? h->t_count : TIMING_RING;
  double sum = 0.0;
  for (int64_t k = 0; k < n; k++) {
    float ms = 0.0f;
    CUDA_TRY(cudaEventElapsedTime(&ms, h->t_ev[2 * k], h->t_ev[2 * k + 1]));
    sum += ms;
  }
  *total_ms = sum; *count = n;
  return PMG_OK;
}

int pmg_debug_box_box(const float* in_host, int64_t n, int32_t stat, float* out_host, int32_t device) {
  if (!in_host || !out_host || n < 1 || stat < 0 || stat > 2) return fail(PMG_ERR_INVALID, "pmg_debug_box_box: invalid argument%s");
  CUDA_TRY(cudaSetDevice(device));
  float *d_in = nullptr, *d_out = nullptr;
  CUDA_TRY(cudaMalloc(&d_in, sizeof(float) * 30 * n));
  CUDA_TRY(cudaMalloc(&d_out, sizeof(float) * 32 * n));
  CUDA_TRY(cudaMemcpy(d_in, in_host, sizeof(float) * 30 * n, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemset(d_out, 0, sizeof(float) * 32 * n));
  debug_box_box_kernel<<<(unsigned)((n + 63) / 64), 64>>>(d_in, n, stat, d_out);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpy(out_host, d_out, sizeof(float) * 32 * n, cudaMemcpyDeviceToHost));
  cudaFree(d_in); cudaFree(d_out);
  return PMG_OK;
}

#ifdef PMG_COOP_TIMING
// development builds only (tools/coop_timing.py): read and clear the cycle counters of pmg_coop.cuh
int pmg_debug_coop_cycles(unsigned long long* out16) {
  cudaDeviceSynchronize();
  unsigned long long zero[16] = {0};
  if (cudaMemcpyFromSymbol(out16, pmg::g_coop_cycles, sizeof zero) != cudaSuccess) return -2;
  return cudaMemcpyToSymbol(pmg::g_coop_cycles, zero, sizeof zero) == cudaSuccess ? 0 : -2;
}
#endif

int pmg_action_error(pmg_handle* h, int32_t clear) {
  if (!h) return 0;
  const int v = *(volatile int*)h->bad_host;
  if (clear) *(volatile int*)h->bad_host = 0;
  return v != 0;
}

int64_t pmg_overflow_count(pmg_handle* h) {
  if (!h) return 0;
  int v = 0;
  cudaSetDevice(h->cfg.device);
  cudaDeviceSynchronize();
  cudaMemcpy(&v, h->d_overflow, sizeof(int), cudaMemcpyDeviceToHost);
  return v;
}

}  // extern "C"
