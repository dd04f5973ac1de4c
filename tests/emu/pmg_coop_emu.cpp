// pmg_coop_emu.cpp -- TEST INFRASTRUCTURE: runs the lane-cooperative device code of
// pybullet_multigoal_gym_b200/csrc/pmg_coop.cuh on the CPU, unchanged.
//
// The 8 lanes of one environment are coroutines (ucontext) scheduled round-robin; every group
// primitive (shuffle, ballot, sync) is a rendezvous: a lane publishes its operand, yields, and reads
// the others' operands when it is resumed, i.e. exactly the lockstep semantics of __shfl_sync /
// __syncwarp over the octet's mask.  A lane that executes a different number of rendezvous than its
// siblings (a divergent collective, which would deadlock or corrupt on the GPU) aborts the run.
// Built by tests/emu/build_emu.py into tests/emu/libpmg_coop_emu.so; used by tests/test_coop_emu.py
// to compare the cooperative algorithm with the CPU oracle without a GPU.  Never loaded by the product.
#define PMG_EMULATE 1
#include <stdio.h>
#include <stdlib.h>
#include <ucontext.h>

#include "../../pybullet_multigoal_gym_b200/csrc/pmg_coop.cuh"
#include "../../pybullet_multigoal_gym_b200/csrc/pmg_spawn.cuh"

namespace pmg_emu {

constexpr int NL = 8;
constexpr size_t STACK = 1 << 20;

struct Sched {
  ucontext_t main_ctx, ctx[NL];
  char* stack[NL];
  bool done[NL];
  long nsync[NL];
  float slot[NL];
  bool pslot[NL];
  int cur;
  void (*body)(int lane, void* arg);
  void* arg;
};
static thread_local Sched* S = nullptr;

static void yield_() {
  S->nsync[S->cur]++;
  swapcontext(&S->ctx[S->cur], &S->main_ctx);
}
int lane() { return S->cur; }
float shfl(float v, int src) {
  S->slot[S->cur] = v;
  yield_();
  float r = S->slot[src & (NL - 1)];
  yield_();
  return r;
}
unsigned ballot(bool p) {
  S->pslot[S->cur] = p;
  yield_();
  unsigned m = 0;
  for (int i = 0; i < NL; i++) m |= (S->pslot[i] ? 1u : 0u) << i;
  yield_();
  return m;
}
void sync() { yield_(); }

static void trampoline(int lane_id) {
  S->body(lane_id, S->arg);
  S->done[lane_id] = true;
  swapcontext(&S->ctx[lane_id], &S->main_ctx);
}

// run body(lane, arg) on 8 lockstep lanes; returns 0, or -1 when the lanes' collectives diverged
static int run_group(void (*body)(int, void*), void* arg) {
  Sched sch;
  S = &sch;
  sch.body = body; sch.arg = arg;
  for (int i = 0; i < NL; i++) {
    sch.stack[i] = (char*)malloc(STACK);
    sch.done[i] = false; sch.nsync[i] = 0;
    getcontext(&sch.ctx[i]);
    sch.ctx[i].uc_stack.ss_sp = sch.stack[i];
    sch.ctx[i].uc_stack.ss_size = STACK;
    sch.ctx[i].uc_link = &sch.main_ctx;
    makecontext(&sch.ctx[i], (void (*)())trampoline, 1, i);
  }
  int rc = 0;
  for (;;) {
    int alive = 0;
    for (int i = 0; i < NL; i++) {
      if (sch.done[i]) continue;
      sch.cur = i;
      swapcontext(&sch.main_ctx, &sch.ctx[i]);
      if (!sch.done[i]) alive++;
    }
    if (alive == 0) break;
    if (alive != NL) {  // some lanes finished while others still wait at a rendezvous
      bool all_same = true;
      for (int i = 1; i < NL; i++) if (sch.done[i] != sch.done[0]) all_same = false;
      if (!all_same) { fprintf(stderr, "pmg_coop_emu: divergent collective (lanes finished at different rendezvous counts)\n"); rc = -1; break; }
    }
  }
  for (int i = 1; i < NL; i++) if (sch.nsync[i] != sch.nsync[0]) { if (rc == 0) fprintf(stderr, "pmg_coop_emu: rendezvous count mismatch lane %d: %ld vs %ld\n", i, sch.nsync[i], sch.nsync[0]); rc = -1; }
  for (int i = 0; i < NL; i++) free(sch.stack[i]);
  S = nullptr;
  return rc;
}

}  // namespace pmg_emu

using namespace pmg;

namespace {
const float* lane_table() {
  static float tab[coop::GL * coop::LC_W];
  static bool init = false;
  if (!init) { for (int l = 0; l < coop::GL; l++) coop::fill_lane_constants(tab + l * coop::LC_W, l); init = true; }
  return tab;
}
struct StepArgs { coop::EnvSmem* sm; StepIO io; };
void step_body(int lane, void* arg) {
  StepArgs* a = (StepArgs*)arg;
  coop::Grp g; g.lane = lane;
  coop::step_env_reach<false>(g, *a->sm, lane_table(), a->io, 0);
}

struct MinvArgs { coop::EnvSmem* sm; const float* q; const float* qd; float* minv_out; float* q_out; float* qd_out; };
void substep_body(int lane, void* arg) {
  MinvArgs* a = (MinvArgs*)arg;
  coop::Grp g; g.lane = lane;
  coop::Lane L;
  L.lc = lane_table() + lane * coop::LC_W; L.dof0 = lane;
  L.q0 = a->q[lane]; L.qd0 = a->qd[lane];
  L.q1 = lane == 7 ? a->q[8] : 0.0f; L.qd1 = lane == 7 ? a->qd[8] : 0.0f;
  L.mt0 = L.q0; L.mt1 = L.q1; L.mi0 = L.mi1 = 0.0f;  // motors off
  L.dtau0 = L.dtau1 = 0.0f;
  coop::substep(g, *a->sm, L);
  a->q_out[lane] = L.q0; a->qd_out[lane] = L.qd0;
  if (lane == 7) { a->q_out[8] = L.q1; a->qd_out[8] = L.qd1; }
}
}  // namespace

extern "C" {

// One env.step() of a Reach environment.  state: the Reach state words (Dims<0,0>::STATE = 50) of one
// env; manifold: 2 x 41 words; both updated in place.  obs_row: 12 floats.
int pmg_emu_reach_step(float* state, float* manifold, const float* action, float thr, int binary, int max_steps,
                       float* obs_row, float* reward, uint8_t* done, uint8_t* success) {
  static coop::EnvSmem sm;
  memset(&sm, 0, sizeof sm);
  StepArgs a;
  a.sm = &sm;
  memset(&a.io, 0, sizeof a.io);
  a.io.state = state; a.io.manifold = manifold; a.io.batch = 1; a.io.tile = 1; a.io.man_words = 0; a.io.state_words = Dims<0, 0>::STATE;
  a.io.action = action; a.io.obs = obs_row; a.io.reward = reward; a.io.done = done; a.io.success = success;
  a.io.thr = thr; a.io.binary = binary; a.io.max_steps = max_steps; a.io.overflow = nullptr; a.io.epw = 4;
  return pmg_emu::run_group(step_body, &a);
}

// One free substep (motors off, no manifolds) from (q, qd); also returns M^-1 (9x9) for cross-checks.
int pmg_emu_substep(const float* q, const float* qd, float* minv81, float* q_out, float* qd_out) {
  static coop::EnvSmem sm;
  memset(&sm, 0, sizeof sm);
  MinvArgs a{&sm, q, qd, minv81, q_out, qd_out};
  int rc = pmg_emu::run_group(substep_body, &a);
  for (int r = 0; r < 9; r++) for (int c = 0; c < 9; c++) minv81[r * 9 + c] = sm.minv[r * coop::MINV_LD + c];
  return rc;
}

int pmg_emu_state_words(void) { return Dims<0, 0>::STATE; }
int pmg_emu_smem_bytes(void) { return (int)sizeof(coop::EnvSmem); }
int pmg_emu_table_bytes(void) { return (int)(coop::GL * coop::LC_W * sizeof(float)); }

}  // extern "C"

// ---- the thread-per-env device code (pmg_sim.cuh) on the host: plain scalar code, no lanes needed ----------
namespace {
template <int NBLK>
void thread_substeps(float* st, float* man, int n_calls) {
  Env<NBLK> e;
  for (int k = 0; k < ND; k++) { e.q[k] = st[ST_Q + k]; e.qd[k] = st[ST_QD + k]; e.mt[k] = st[ST_MT + k]; e.mi[k] = st[ST_MI + k]; e.dtau[k] = 0; }
  for (int b = 0; b < NBLK; b++) {
    const float* bs = st + ST_BLK + 13 * b;
    e.bpos[b] = v3(bs[0], bs[1], bs[2]);
    for (int k = 0; k < 4; k++) e.bquat[b][k] = bs[3 + k];
    e.bv[b] = v3(bs[7], bs[8], bs[9]); e.bw[b] = v3(bs[10], bs[11], bs[12]);
  }
  e.man = man; e.stride = 1; e.overflow = 0;
  for (int c = 0; c < n_calls; c++) {
    for (int k = 0; k < ND; k++) e.dtau[k] = -c_dof_damping[k] * e.qd[k];
    for (int s = 0; s < SUBSTEPS_PER_CALL; s++) substep(e);
  }
  for (int k = 0; k < ND; k++) { st[ST_Q + k] = e.q[k]; st[ST_QD + k] = e.qd[k]; }
  for (int b = 0; b < NBLK; b++) {
    float* bs = st + ST_BLK + 13 * b;
    bs[0] = e.bpos[b].x; bs[1] = e.bpos[b].y; bs[2] = e.bpos[b].z;
    for (int k = 0; k < 4; k++) bs[3 + k] = e.bquat[b][k];
    bs[7] = e.bv[b].x; bs[8] = e.bv[b].y; bs[9] = e.bv[b].z; bs[10] = e.bw[b].x; bs[11] = e.bw[b].y; bs[12] = e.bw[b].z;
  }
}
}  // namespace

extern "C" {
// n_calls x 20 substeps of the thread-per-env kernel's physics on one env: state = the usual state words
// (q qd ee rest mt mi | blocks ...), manifold = num_pairs(nblk) x 41 words, both updated in place.
int pmg_emu_thread_substeps(int nblk, float* state, float* manifold, int n_calls) {
  switch (nblk) {
    case 0: thread_substeps<0>(state, manifold, n_calls); return 0;
    case 1: thread_substeps<1>(state, manifold, n_calls); return 0;
    case 2: thread_substeps<2>(state, manifold, n_calls); return 0;
    case 3: thread_substeps<3>(state, manifold, n_calls); return 0;
    default: return -1;
  }
}
}

// ---- the one-block cooperative step (Push / PickAndPlace) ------------------------------------------------------
namespace {
struct BlkArgs { coop::EnvSmemT<1>* sm; coop::EnvSmemT<1, true>* smp; StepIO io; int task; };
void blk_body(int lane, void* arg) {
  BlkArgs* a = (BlkArgs*)arg;
  coop::Grp g; g.lane = lane;
  if (a->task == 1) coop::step_env_block<1>(g, *a->sm, lane_table(), a->io, 0);
  else if (a->task == 5) coop::step_env_block<5>(g, *a->smp, lane_table(), a->io, 0);   // Slide: long table + puck
  else coop::step_env_block<2>(g, *a->sm, lane_table(), a->io, 0);
}
}  // namespace

extern "C" {
// One env.step() of a Push (task 1, 3 action columns) / PickAndPlace (task 2, 4 columns) environment.  state:
// Dims<TASK,1>::STATE = 63 words of one env; manifold: 6 x 41 words; obs_row: 33 floats.
int pmg_emu_block_step(int task, float* state, float* manifold, const float* action, float thr, int binary, int max_steps,
                       float* obs_row, float* reward, uint8_t* done, uint8_t* success) {
  static coop::EnvSmemT<1> sm;
  static coop::EnvSmemT<1, true> smp;
  memset(&sm, 0, sizeof sm);
  memset(&smp, 0, sizeof smp);
  BlkArgs a;
  a.sm = &sm; a.smp = &smp; a.task = task;
  memset(&a.io, 0, sizeof a.io);
  a.io.state = state; a.io.manifold = manifold; a.io.batch = 1; a.io.tile = 1; a.io.man_words = 0; a.io.state_words = Dims<1, 1>::STATE;
  a.io.action = action; a.io.obs = obs_row; a.io.reward = reward; a.io.done = done; a.io.success = success;
  a.io.thr = thr; a.io.binary = binary; a.io.max_steps = max_steps; a.io.overflow = nullptr; a.io.epw = 4;
  a.io.grasp = task == 2; a.io.adim = task == 2 ? 4 : 3; a.io.goal_dim = 3; a.io.row_width = 33;
  static float spill[coop::EnvSmemT<1>::SPILL_WORDS];
  a.io.row_spill = spill;
  return pmg_emu::run_group(blk_body, &a);
}
int pmg_emu_block_smem_bytes(void) { return (int)sizeof(coop::EnvSmemT<1>); }

}  // extern "C"

// One env.step() of a multi-block environment (BlockStack layout, nb = 2..5 blocks, 4 action columns): state
// (state_words of one env) and manifolds (num_pairs(nb) x 41 words) updated in place, packed row out.
// grip: 1 = grip-informed goal (goal_dim 3 nb + 4), -1 = BlockRearrange (no grasping, 3 action columns);
// td: task decomposition / curriculum (sub-goal word in the state).
template <int NB>
static int multi_step(float* state, float* manifold, const float* action, int* overflow, int grip, int td, float thr, int binary,
                      int max_steps, float* obs_row, float* reward, uint8_t* done, uint8_t* success) {
  static coop::EnvSmemT<NB> sm;
  static float spill[coop::EnvSmemT<NB>::SPILL_WORDS];
  memset(&sm, 0, sizeof sm);
  struct Args { coop::EnvSmemT<NB>* sm; StepIO io; } a;
  a.sm = &sm;
  memset(&a.io, 0, sizeof a.io);
  const int G = 3 * NB + (grip > 0 ? 4 : 0);
  a.io.state = state; a.io.manifold = manifold; a.io.batch = 1; a.io.tile = 1; a.io.man_words = 0; a.io.state_words = ST_BLK + 13 * NB + G + (td ? 1 : 0) + 1;
  a.io.action = action; a.io.obs = obs_row; a.io.reward = reward; a.io.done = done; a.io.success = success;
  a.io.thr = thr; a.io.binary = binary; a.io.max_steps = max_steps;
  a.io.overflow = overflow; a.io.epw = 4; a.io.grasp = grip >= 0; a.io.adim = grip >= 0 ? 4 : 3; a.io.row_spill = spill;
  if (grip < 0) grip = 0;  // grip = -1: BlockRearrange (no grasping, 3 action columns)
  a.io.grip_goal = grip; a.io.td = td; a.io.goal_dim = G; a.io.row_width = Dims<3, NB>::O + Dims<3, NB>::P + 2 * G;
  return pmg_emu::run_group([](int lane, void* arg) {
    Args* a = (Args*)arg;
    coop::Grp g; g.lane = lane;
    coop::step_env_multi<NB>(g, *a->sm, lane_table(), a->io, 0);
  }, &a);
}
extern "C" {
int pmg_emu_multi_step(int nb, float* state, float* manifold, const float* action, int* overflow, int grip, int td, float thr,
                       int binary, int max_steps, float* obs_row, float* reward, uint8_t* done, uint8_t* success) {
  switch (nb) {
    case 2: return multi_step<2>(state, manifold, action, overflow, grip, td, thr, binary, max_steps, obs_row, reward, done, success);
    case 3: return multi_step<3>(state, manifold, action, overflow, grip, td, thr, binary, max_steps, obs_row, reward, done, success);
    case 4: return multi_step<4>(state, manifold, action, overflow, grip, td, thr, binary, max_steps, obs_row, reward, done, success);
    case 5: return multi_step<5>(state, manifold, action, overflow, grip, td, thr, binary, max_steps, obs_row, reward, done, success);
  }
  return -1;
}
int pmg_emu_multi_smem_bytes(int nb) {
  return nb == 2 ? (int)sizeof(coop::EnvSmemT<2>) : nb == 3 ? (int)sizeof(coop::EnvSmemT<3>) : nb == 4 ? (int)sizeof(coop::EnvSmemT<4>) : (int)sizeof(coop::EnvSmemT<5>);
}

// The joint-control variants (kuka.py:204-206): task 0 Reach (7 action columns, row of 26 floats), 1 Push (7, 47),
// 2 PickAndPlace (8, 47).
int pmg_emu_step_jc(int task, float* state, float* manifold, const float* action, float thr, int binary, int max_steps,
                    float* obs_row, float* reward, uint8_t* done, uint8_t* success) {
  StepIO io;
  memset(&io, 0, sizeof io);
  io.state = state; io.manifold = manifold; io.batch = 1; io.tile = 1; io.man_words = 0; io.state_words = task == 0 ? Dims<0, 0>::STATE : Dims<1, 1>::STATE;
  io.action = action; io.obs = obs_row; io.reward = reward; io.done = done; io.success = success;
  io.thr = thr; io.binary = binary; io.max_steps = max_steps; io.overflow = nullptr; io.epw = 4;
  io.jc = 1; io.grasp = task == 2; io.adim = task == 2 ? 8 : 7; io.goal_dim = 3; io.row_width = task == 0 ? 26 : 47;
  if (task == 0) {
    static coop::EnvSmem sm;
    memset(&sm, 0, sizeof sm);
    StepArgs a; a.sm = &sm; a.io = io;
    return pmg_emu::run_group([](int lane, void* arg) {
      StepArgs* a = (StepArgs*)arg;
      coop::Grp g; g.lane = lane;
      coop::step_env_reach<true>(g, *a->sm, lane_table(), a->io, 0);
    }, &a);
  }
  static coop::EnvSmemT<1> sm;
  static float spill[coop::EnvSmemT<1>::SPILL_WORDS];
  memset(&sm, 0, sizeof sm);
  io.row_spill = spill;
  BlkArgs a; a.sm = &sm; a.io = io; a.task = task;
  return pmg_emu::run_group(blk_body, &a);
}
}

// the device-side reset sampler (csrc/pmg_spawn.cuh) on the host, for the bit-exact comparison with
// oracle/device_rng_oracle.py in the CPU test suite
extern "C" void pmg_emu_device_spawn(int task, int nb, int grip, unsigned long long seed, long long env, unsigned episode, float* out) {
  double tip[3], ol[3], oh[3], tl[3], th[3];
  pmg::spawn::task_bounds(task, tip, ol, oh, tl, th);
  const pmg::spawn::Bounds b = pmg::spawn::to_bounds(tip, ol, oh, tl, th);
  pmg::spawn::Philox r;
  pmg::spawn::stream_init(r, seed, env, episode);
  pmg::spawn::sample_row(r, task, nb, grip, b, out);
}
extern "C" void pmg_emu_philox(const unsigned* ctr, const unsigned* key, unsigned* out) { pmg::spawn::philox_block(ctr, key, out); }

// box_box on one pair of boxes: in = p1[3] R1[9, rows] A[3] p2[3] R2[9] B[3]; out = count, then 4 x (pB[3] nB[3] dist).
// stat selects the static fast path (pmg_physics.cuh); the equivalence test runs every case with stat = 0 as well.
extern "C" int pmg_emu_box_box(const float* in, int stat, float* out) {
  using namespace pmg;
  auto m3 = [](const float* r) { M3 m; m.r0 = v3(r[0], r[1], r[2]); m.r1 = v3(r[3], r[4], r[5]); m.r2 = v3(r[6], r[7], r[8]); return m; };
  BoxScratch scr;
  const M3 R1 = m3(in + 3), R2 = m3(in + 18);
  const int n = box_box(v3(in[0], in[1], in[2]), R1, v3(in[12], in[13], in[14]), v3(in[15], in[16], in[17]), R2, v3(in[27], in[28], in[29]), scr, stat);
  out[0] = (float)n;
  for (int i = 0; i < n; i++) {
    float* o = out + 1 + 7 * i;
    o[0] = scr.out[i].pB.x; o[1] = scr.out[i].pB.y; o[2] = scr.out[i].pB.z;
    o[3] = scr.out[i].nB.x; o[4] = scr.out[i].nB.y; o[5] = scr.out[i].nB.z; o[6] = scr.out[i].dist;
  }
  return n;
}
