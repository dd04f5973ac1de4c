// pmg_coop.cuh -- lane-cooperative env.step(): 8 lanes of a warp own ONE environment.
//
// Why: at the reference batch (8192 envs) a thread-per-env kernel is 256 warps on 592 warp schedulers,
// each warp grinding through ~10 K dependent instructions per substep with its working set spilled to
// L1-resident local memory (profiles/r01_*).  Here an environment is spread over the 8 lanes of an
// "octet" (4 environments per warp, 2048 warps at batch 8192 = 3.5 warps per scheduler):
//   * lane c = chain body c (link_1..link_7, gripper base); the link frames are an inclusive SCAN of
//     rigid transforms over the lanes (3 shuffle levels instead of a 8-step serial recursion), the
//     velocity / acceleration recursions are prefix sums, the composite-rigid-body inertias and the
//     subtree wrenches are suffix sums (all about one common reference point, the gripper-base
//     origin, so that they ARE plain sums);
//   * the two fingers hang off the gripper base with the same orientation and are evaluated
//     redundantly by every lane (cheaper than exchanging them);
//   * lane r owns row r of the 9x9 joint-space inertia matrix (lane 7 owns the two finger rows); the
//     inverse is an in-place Gauss-Jordan sweep with one row broadcast per pivot;
//   * projected Gauss-Seidel: the delta-velocity vector is distributed over the lanes, the owner lane
//     of a motor / limit row computes the impulse and broadcasts it (one shuffle per row); contact
//     rows are built with lanes over contact points (J, M^-1 J^T records in shared memory) and swept
//     replicated on a gathered delta-velocity vector (no shuffles on a path that usually runs diverged);
//   * narrowphase: lanes over collision pairs; Push / PickAndPlace (EnvSmemT<1>) add the block body, four
//     more pairs and rows with a block end point, specialised by which ends a row has;
//   * no per-thread work arrays in local memory: per-lane state is ~100 registers, exchange goes through shuffles and
//     3.6 KB (Reach) / 7.1 KB (one block) / 13.4 KB (four blocks) of shared memory per environment (what remains in
//     local memory are ptxas' register spills: ~100 bytes per thread in the Reach kernel at its 128-register budget,
//     none in the one-block kernels, ~40 bytes in the multi-block kernels);
//   * multi-block scenes (NBLK = 2..5): lane b owns block b and solves its static (table / floor) rows on its own.
// The arithmetic is the same system the thread-per-env kernel (pmg_sim.cuh) and the oracle solve --
// reference call sequence robots/kuka.py:167-225, envs/base_envs/base_env.py:215-219 -- organised for
// lanes; the group primitives below are the only device-specific part, and tests/emu/ runs this very
// source on the CPU (PMG_EMULATE) against the oracle.
#pragma once

#include <type_traits>

#include "pmg_sim.cuh"

namespace pmg {
namespace coop {

constexpr int GL = 8;        // lanes per environment
#ifndef PMG_SPTS1
#define PMG_SPTS1 12         // one-block kernels: contact points whose rows live in shared memory (the rest: global spill); 11 measured slower, DESIGN.md section 9
#endif
constexpr int R_J = 0, R_RHS = 9, R_DINV = 10, R_MJ = 12, R_DENOM = 21, R_MU = 22;  // contact row record, robot part
constexpr int R_BD = 24, R_BA = 28;  // block part (NBLK > 0): direction d[3] has_robot | lever arm r_B x d [3] has_block
constexpr int R_AA = 32, R_IDX = 36;  // NBLK > 1: lever arm r_A x d [3] . | block index of end A, of end B (-1: none)
constexpr int MINV_LD = 12;  // row stride of M^-1 in shared memory

// Shared memory of one environment.  NBLK = 0: Reach (two finger-table pairs, <= 8 contact points);
// NBLK = 1: Push / PickAndPlace (+ table-block, floor-block, finger-block pairs: 6 pairs x 4 points, the same pool
// as the thread-per-env kernels; a grasp on the table really uses 20 of them).  Only the records of the first
// SPTS points live in shared memory; the rarely used rest goes to a per-environment global scratch (`spill`, L1/L2
// resident, read as warp-uniform broadcasts), which is what lets 7 blocks = 28 environments share one SM.
template <int NBLK> struct RowSpill { float* spill; float spill_pad_[2]; };
template <> struct RowSpill<0> {};
// NBLK > 1: the points that are not static, in pair order (point index into the impulse arrays), built once per substep
// (glist: point index into the impulse arrays), with each one's pair and its index inside the pair's manifold
template <int NBLK> struct GenList { unsigned char glist[48], ppair[48], pidx[48]; };
template <> struct GenList<0> {};
template <> struct GenList<1> {};
template <int NBLK, bool PUCK_ = false>
struct __align__(16) EnvSmemT : RowSpill<NBLK>, GenList<NBLK> {
  static constexpr int NB = NBLK;
  static constexpr bool PUCK = PUCK_;  // Slide: long table, the block is a cylinder with an anisotropic inertia
  static constexpr int NPAIRS = num_pairs(NBLK);                // 2 / 6
  static constexpr int MAXPTS = NBLK == 0 ? 8 : (NBLK == 1 ? 24 : 48);  // cached contact points that get rows
  static constexpr int SPTS = NBLK == 0 ? 8 : (NBLK == 1 ? PMG_SPTS1 : 8);  // ... of which in shared memory (NBLK > 1: of the points that are not static)
  static constexpr int ROW_W = NBLK == 0 ? 24 : (NBLK == 1 ? 32 : 40);  // floats per contact row record (16-byte aligned parts)
  static constexpr int SROWS = NBLK > 1 ? NBLK * 4 * 3 : 0;     // NBLK > 1: compact rows of the static points (table / floor against a block), 4 points per block
  static constexpr int SROW_W = 12;                             // d[3] rhs | r_B x d [3] dinv | denom mu . .
  static constexpr int SROWS_AS_ROWS = (SROWS * SROW_W + ROW_W - 1) / ROW_W;  // ... stored behind the general rows
  static constexpr int NSCR = NPAIRS < GL ? NPAIRS : GL;        // narrowphase scratch areas (one per lane that runs pairs)
  float pub[7][8];                  // per arm dof: axis a, v = (p - Pref) x a
  float minv[ND * MINV_LD];         // 9x9, rows padded to 12
  float man[(NPAIRS * MAN_WORDS + 3) / 4 * 4];  // persistent manifolds (41 words per pair)
  float hand[24];                   // gripper frame for the contact rows: Rg[9] pf1 pf2 Pref ax1 (kept out of registers)
  float vq[NBLK == 0 ? 12 : (ND + 6 * NBLK + 3) / 4 * 4];  // joint (+ block) velocities for the row set-up, then the PGS delta velocities
  float rows[SPTS * 3 + SROWS_AS_ROWS][ROW_W];  // contact rows (layout R_* above), then the compact static rows; narrowphase scratch before they are built
  float app[2][MAXPTS * 3];         // accumulated impulses of the contact rows, double buffered over the PGS iterations
  float blk[NBLK > 0 ? 24 * NBLK : 1];  // per block: pos[3] quat[4] v[3] w[3] R[9] (+ 2 spare words)
  static constexpr int SCRATCH_FLOATS = (SPTS * 3 + SROWS_AS_ROWS) * ROW_W;
  __device__ __forceinline__ float* srow(int r) { return &rows[0][0] + SPTS * 3 * ROW_W + r * SROW_W; }  // compact static row r
  static_assert(NSCR * sizeof(BoxScratch) <= SCRATCH_FLOATS * sizeof(float), "narrowphase scratch must fit in the contact-row area");
  static constexpr int SPILL_WORDS = (MAXPTS - SPTS) * 3 * ROW_W;  // global scratch per environment
  __device__ __forceinline__ float* spill_row(int r) { return this->spill + (r - SPTS * 3) * ROW_W; }  // r >= 3 SPTS
};
using EnvSmem = EnvSmemT<0>;
// Distance between the shared-memory copies of the environments of one block.  The four octets of a warp execute every
// shared-memory instruction together, each on its own copy at the same offset; sizeof(EnvSmemT<0>) and <1> are 16 banks
// mod 32, which puts octets 0 / 2 and 1 / 3 on the same banks: every LDS / STS of the kernel a 2-way conflict
// (ncu: 124 M conflict wavefronts per Push launch).  A stride of 24 banks (96 bytes mod 128) lays the octets' banks
// side by side (0, 24, 16, 8): conflict-free for every access in which an octet touches <= 8 consecutive banks.
template <class SM>
__host__ __device__ constexpr size_t env_stride() {
#ifdef PMG_ENV_STRIDE_PLAIN
  return sizeof(SM);
#else
  size_t s = sizeof(SM);
  while (s % 128 != 96) s += 16;
  return s;
#endif
}
constexpr int BK_POS = 0, BK_QUAT = 3, BK_V = 7, BK_W = 10, BK_R = 13;

// ---- the group (octet) interface ---------------------------------------------------------------
struct Grp {
  int lane;  // 0..7 inside the environment
#ifdef PMG_EMULATE
  __device__ float shfl(float v, int src) const { return pmg_emu::shfl(v, src); }
  __device__ unsigned ballot(bool p) const { return pmg_emu::ballot(p); }
  __device__ unsigned reduce_or(unsigned v) const {  // bitwise OR over the octet
    unsigned r = 0;
    for (int b = 0; b < 32; b++) if (pmg_emu::ballot(((v >> b) & 1u) != 0)) r |= 1u << b;
    return r;
  }
  __device__ void sync() const { pmg_emu::sync(); }
  __device__ void block_sync() const {}
#else
  unsigned mask;  // the octet's lanes inside the warp
  int shift;      // first lane of the octet
  // Warps of one block that run the substep loop side by side share its instruction stream: the loop body (39 KB
  // without contacts, 90-130 KB with) does not fit the 32 KB instruction cache, so a warp on its own streams it
  // from L2 every substep; the first of a group that stays together fetches a line, the others hit.
  // (blocks of several warps meet at the top of every substep; blockDim sits in the constant bank: no register)
  __device__ __forceinline__ void block_sync() const { if (blockDim.x > 32) __syncthreads(); }
  __device__ __forceinline__ float shfl(float v, int src) const { return __shfl_sync(mask, v, src, GL); }
  __device__ __forceinline__ unsigned ballot(bool p) const { return (__ballot_sync(mask, p) >> shift) & 0xffu; }
  __device__ __forceinline__ unsigned reduce_or(unsigned v) const { return __reduce_or_sync(mask, v); }  // one REDUX
  __device__ __forceinline__ void sync() const { __syncwarp(mask); }
#endif
  // value of lane - d (own value when there is no such lane) / lane + d / lane ^ m
#ifdef PMG_EMULATE
  __device__ __forceinline__ float up(float v, int d) const { return shfl(v, lane >= d ? lane - d : lane); }
  __device__ __forceinline__ float down(float v, int d) const { return shfl(v, lane + d < GL ? lane + d : lane); }
  __device__ __forceinline__ float bfly(float v, int m) const { return shfl(v, lane ^ m); }
#else  // the native forms clamp inside the 8-lane segment themselves: no source-lane arithmetic
  __device__ __forceinline__ float up(float v, int d) const { return __shfl_up_sync(mask, v, d, GL); }
  __device__ __forceinline__ float down(float v, int d) const { return __shfl_down_sync(mask, v, d, GL); }
  __device__ __forceinline__ float bfly(float v, int m) const { return __shfl_xor_sync(mask, v, m, GL); }
#endif
  __device__ __forceinline__ V3 shfl(V3 v, int src) const { return v3(shfl(v.x, src), shfl(v.y, src), shfl(v.z, src)); }
  __device__ __forceinline__ V3 up(V3 v, int d) const { return v3(up(v.x, d), up(v.y, d), up(v.z, d)); }
  __device__ __forceinline__ V3 down(V3 v, int d) const { return v3(down(v.x, d), down(v.y, d), down(v.z, d)); }
  __device__ __forceinline__ M3 shfl(const M3& m, int src) const { M3 r; r.r0 = shfl(m.r0, src); r.r1 = shfl(m.r1, src); r.r2 = shfl(m.r2, src); return r; }
  __device__ __forceinline__ M3 up(const M3& m, int d) const { M3 r; r.r0 = up(m.r0, d); r.r1 = up(m.r1, d); r.r2 = up(m.r2, d); return r; }
  __device__ __forceinline__ float sum(float v) const { v += bfly(v, 4); v += bfly(v, 2); v += bfly(v, 1); return v; }
  __device__ __forceinline__ float maxv(float v) const { v = fmaxf(v, bfly(v, 4)); v = fmaxf(v, bfly(v, 2)); v = fmaxf(v, bfly(v, 1)); return v; }
  __device__ __forceinline__ V3 scan(V3 v) const {  // inclusive prefix sum over the lanes
#pragma unroll
    for (int d = 1; d < GL; d <<= 1) { V3 o = up(v, d); if (lane >= d) v += o; }
    return v;
  }
  __device__ __forceinline__ float rscan(float v) const {  // inclusive suffix sum
#pragma unroll
    for (int d = 1; d < GL; d <<= 1) { float o = down(v, d); if (lane + d < GL) v += o; }
    return v;
  }
  __device__ __forceinline__ V3 rscan(V3 v) const { return v3(rscan(v.x), rscan(v.y), rscan(v.z)); }
};

// ---- per-lane registers ----------------------------------------------------------------------------
// Dof slots: slot 0 is arm joint `lane` on lanes 0..6 and finger1 (dof 7) on lane 7; slot 1 is finger2
// (dof 8) on lane 7 and unused (zero) elsewhere.
struct Lane {
  const float* lc;  // this lane's row of the constant table (shared memory), see LC_* below
  float q0, q1, qd0, qd1, mt0, mt1, mi0, mi1, dtau0, dtau1;
  int dof0;
};

// Per-lane model constants of chain body `lane` (link_1..link_7, gripper base): one table per block in
// shared memory, read where they are used instead of being held in ~25 registers through the whole substep
// (the kernel has 128 registers per thread: 4 warps x 32 x 128 fill a scheduler's register file).
constexpr int LC_JROT = 0, LC_JXYZ = 9, LC_COM = 12, LC_INERTIA = 15, LC_MASS = 18, LC_MSUB = 19, LC_LOWER = 20, LC_UPPER = 21, LC_DAMP = 22;
constexpr int LC_W = 28;  // row stride: lanes 0..7 hit 8 different banks
__device__ __forceinline__ V3 lc3(const float* lc, int o) { return v3(lc[o], lc[o + 1], lc[o + 2]); }

__device__ inline void fill_lane_constants(float* row, int b) {  // b = chain body = lane
#pragma unroll
  for (int k = 0; k < 9; k++) row[LC_JROT + k] = c_jrot[b][k];
#pragma unroll
  for (int k = 0; k < 3; k++) { row[LC_JXYZ + k] = c_jxyz[b][k]; row[LC_COM + k] = c_com[b][k]; row[LC_INERTIA + k] = c_inertia[b][k]; }
  row[LC_MASS] = c_mass[b];
  float msub = c_mass[PMG_BODY_FINGER1] + c_mass[PMG_BODY_FINGER2];  // mass of the subtree rooted at this body
  for (int k = 7; k >= b; k--) msub += c_mass[k];
  row[LC_MSUB] = msub;
  row[LC_LOWER] = c_dof_lower[b]; row[LC_UPPER] = c_dof_upper[b]; row[LC_DAMP] = c_dof_damping[b];
}

__device__ __forceinline__ Sym3 sym_add(const Sym3& a, const Sym3& b) {
  Sym3 r; r.xx = a.xx + b.xx; r.xy = a.xy + b.xy; r.xz = a.xz + b.xz; r.yy = a.yy + b.yy; r.yz = a.yz + b.yz; r.zz = a.zz + b.zz; return r;
}
// R diag(i) R^T
__device__ __forceinline__ Sym3 world_inertia(const M3& R, V3 i) {
  V3 s0 = v3(R.r0.x * i.x, R.r0.y * i.y, R.r0.z * i.z), s1 = v3(R.r1.x * i.x, R.r1.y * i.y, R.r1.z * i.z), s2 = v3(R.r2.x * i.x, R.r2.y * i.y, R.r2.z * i.z);
  Sym3 I;
  I.xx = dot(s0, R.r0); I.xy = dot(s0, R.r1); I.xz = dot(s0, R.r2);
  I.yy = dot(s1, R.r1); I.yz = dot(s1, R.r2); I.zz = dot(s2, R.r2);
  return I;
}
// m (c.c 1 - c c^T)
__device__ __forceinline__ Sym3 point_inertia(float m, V3 c) {
  const float cc = dot(c, c);
  Sym3 I;
  I.xx = m * (cc - c.x * c.x); I.xy = -m * c.x * c.y; I.xz = -m * c.x * c.z;
  I.yy = m * (cc - c.y * c.y); I.yz = -m * c.y * c.z; I.zz = m * (cc - c.z * c.z);
  return I;
}

// Link frames of the chain: inclusive scan of (R, p) under composition.  qarm = joint angle (0 on lane 7).
__device__ __forceinline__ void chain_fk(const Grp& g, const Lane& L, float qarm, M3& R, V3& p) {
  float s, c;
  sincosf(qarm, &s, &c);
  M3 J;  // J * Rz(q): new x column = c*x + s*y, new y column = -s*x + c*y
  J.r0 = lc3(L.lc, LC_JROT); J.r1 = lc3(L.lc, LC_JROT + 3); J.r2 = lc3(L.lc, LC_JROT + 6);
  R.r0 = v3(c * J.r0.x + s * J.r0.y, -s * J.r0.x + c * J.r0.y, J.r0.z);
  R.r1 = v3(c * J.r1.x + s * J.r1.y, -s * J.r1.x + c * J.r1.y, J.r1.z);
  R.r2 = v3(c * J.r2.x + s * J.r2.y, -s * J.r2.x + c * J.r2.y, J.r2.z);
  p = lc3(L.lc, LC_JXYZ);
#pragma unroll
  for (int d = 1; d < GL; d <<= 1) {
    M3 Ro = g.up(R, d);
    V3 po = g.up(p, d);
    if (g.lane >= d) { p = po + mul(Ro, p); R = mul(Ro, R); }
  }
}

// ---- inverse kinematics (same iteration as pmg_physics.cuh::inverse_kinematics) ---------------------
// Lane j owns column j of the tip Jacobian and row j of J^T J + 0.5 I; the 7x7 system is solved by
// Gauss-Jordan elimination with one row broadcast per pivot.  Returns this lane's joint angle.
__device__ float inverse_kinematics(const Grp& g, const Lane& L, float qarm, V3 target, const float tq[4]) {
  const int lane = g.lane;
  const bool arm = lane < 7;
  const float t[3] = PMG_TIP_OFFSET;
  float diff = 1e30f;
  for (int it = 0; it < 40 && diff > 1e-5f; it++) {
    M3 R; V3 p;
    chain_fk(g, L, qarm, R, p);
    M3 R6 = g.shfl(R, PMG_BODY_LINK7);
    V3 p6 = g.shfl(p, PMG_BODY_LINK7);
    V3 tip = p6 + mul(R6, v3(t[0], t[1], t[2]));
    V3 ep = target - tip;
    diff = norm(ep);
    float qc[4];
    m3_to_quat(R6, qc);
    float ax = -qc[0], ay = -qc[1], az = -qc[2], aw = qc[3];
    float dx = tq[3] * ax + tq[0] * aw + tq[1] * az - tq[2] * ay;
    float dy = tq[3] * ay + tq[1] * aw + tq[2] * ax - tq[0] * az;
    float dz = tq[3] * az + tq[2] * aw + tq[0] * ay - tq[1] * ax;
    float dw = tq[3] * aw - tq[0] * ax - tq[1] * ay - tq[2] * az;
    float vn2 = dx * dx + dy * dy + dz * dz;
    V3 er;
    if (vn2 < 10.0f * 2.220446049250313e-16f) {
      float ang = 2.0f * atan2f(sqrtf(vn2), dw);
      if (ang > PI_F) ang -= 2.0f * PI_F;
      er = v3(ang, 0, 0);
    } else {
      float vn = sqrtf(vn2);
      float ang = 2.0f * atan2f(vn, dw);
      if (ang > PI_F) ang -= 2.0f * PI_F;
      float sc = ang / vn;
      er = v3(dx * sc, dy * sc, dz * sc);
    }
    V3 a = arm ? col(R, 2) : v3(0, 0, 0);
    V3 Jl = cross(a, tip - p), Ja = a;
    float A[7];
#pragma unroll
    for (int j = 0; j < 7; j++) {
      V3 Jlj = g.shfl(Jl, j), Jaj = g.shfl(Ja, j);
      A[j] = dot(Jl, Jlj) + dot(Ja, Jaj) + (j == lane ? IK_JOINT_DAMPING : 0.0f);
    }
    float rhs = dot(Jl, ep) + dot(Ja, er);
#pragma unroll
    for (int k = 0; k < 7; k++) {
      float rk[7];
#pragma unroll
      for (int j = k; j < 7; j++) rk[j] = g.shfl(A[j], k);
      const float rr = g.shfl(rhs, k);
      const float pinv = 1.0f / rk[k];
      const bool own = lane == k;
      const float f = own ? -pinv : A[k] * pinv;
#pragma unroll
      for (int j = k + 1; j < 7; j++) A[j] = (own ? 0.0f : A[j]) - f * rk[j];
      rhs = (own ? 0.0f : rhs) - f * rr;
    }
    const float mx = g.maxv(fabsf(rhs));
    const float scale = mx > IK_MAX_STEP ? IK_MAX_STEP / mx : 1.0f;
    if (arm) qarm += scale * rhs;
  }
  return qarm;
}

// ---- solver state of a lane -----------------------------------------------------------------------
struct SolverLane {
  float rhs0, rhs1, lim0, lim1, app0, app1, dinv0, dinv1, mdd0, mdd1;
  float lrhs00, lrhs01, lrhs10, lrhs11;  // [slot][side]
  float lapp00, lapp01, lapp10, lapp11;
  float dqd0, dqd1;
  float big0, big1;  // largest |impulse change| of the lane's own rows in this sweep (convergence test)
};

// Motor row of the K-th visited dof.  Every lane evaluates the row formula on its own slot; only the
// owner's result is kept (predicated moves) and broadcast.
template <int K>
__device__ __forceinline__ void motor_row(const Grp& g, SolverLane& s, const float* A0, const float* A1) {
  constexpr int d = nc_dof(K);
  constexpr int owner = d < 8 ? d : 7;
  constexpr bool slot1 = d == 8;
  const float rhs = slot1 ? s.rhs1 : s.rhs0, cur = slot1 ? s.dqd1 : s.dqd0, dinv = slot1 ? s.dinv1 : s.dinv0;
  const float app = slot1 ? s.app1 : s.app0, lim = slot1 ? s.lim1 : s.lim0;
  float dlc = rhs - cur * dinv;
  const float sum = fminf(fmaxf(app + dlc, -lim), lim);
  dlc = sum - app;
  if (g.lane == owner) {
    if (slot1) { s.app1 = sum; s.big1 = fmaxf(s.big1, fabsf(dlc)); } else { s.app0 = sum; s.big0 = fmaxf(s.big0, fabsf(dlc)); }
  }
  const float dl = g.shfl(dlc, owner);
  s.dqd0 += A0[d] * dl;
  s.dqd1 += A1[d] * dl;
}

// limit row in slot s (visit position s>>1, side s&1); data-dependent => dynamic dof, M^-1 from shared memory
template <class SM>
__device__ __forceinline__ void limit_row(const Grp& g, SolverLane& s, const SM& sm, int slot_id, int dof0) {
  const int d = c_nc_order[ND + (slot_id >> 1)];
  const bool side = (slot_id & 1) != 0, slot1 = d == 8;
  const int owner = d < 8 ? d : 7;
  const float sign = side ? -1.0f : 1.0f;
  const float lr = slot1 ? (side ? s.lrhs11 : s.lrhs10) : (side ? s.lrhs01 : s.lrhs00);
  const float la = slot1 ? (side ? s.lapp11 : s.lapp10) : (side ? s.lapp01 : s.lapp00);
  const float cur = slot1 ? s.dqd1 : s.dqd0, dinv = slot1 ? s.dinv1 : s.dinv0;
  float dlc = lr - sign * cur * dinv;
  const float sum = fminf(fmaxf(la + dlc, 0.0f), LIMIT_MAX_IMPULSE);
  dlc = sum - la;
  if (g.lane == owner) {
    if (slot1) { if (side) s.lapp11 = sum; else s.lapp10 = sum; s.big1 = fmaxf(s.big1, fabsf(dlc)); }
    else { if (side) s.lapp01 = sum; else s.lapp00 = sum; s.big0 = fmaxf(s.big0, fabsf(dlc)); }
  }
  const float sdl = sign * g.shfl(dlc, owner);
  s.dqd0 += sm.minv[d * MINV_LD + dof0] * sdl;
  s.dqd1 += sm.minv[d * MINV_LD + 8] * sdl;
}

// The closed jaws of Reach / Push sit on their upper limit, so the finger limit rows are active in most
// substeps: static versions of exactly those rows (dof and owner known at compile time, M^-1 from registers).
__host__ __device__ constexpr int visit_pos(int dof) {
  for (int k = 0; k < ND; k++) if (nc_dof(k) == dof) return k;
  return -1;
}
constexpr int K_F1 = visit_pos(7), K_F2 = visit_pos(8);
static_assert(K_F1 >= 0 && K_F2 == K_F1 + 1, "finger limit fast path assumes finger1, finger2 are visited back to back");
constexpr unsigned FINGER_LIMIT_BITS = (3u << (2 * K_F1)) | (3u << (2 * K_F2));

template <int DOF, int SIDE>
__device__ __forceinline__ void finger_limit_row(const Grp& g, SolverLane& s, const float* A0, const float* A1) {
  constexpr bool slot1 = DOF == 8;
  constexpr float sign = SIDE ? -1.0f : 1.0f;
  float& lapp = slot1 ? (SIDE ? s.lapp11 : s.lapp10) : (SIDE ? s.lapp01 : s.lapp00);
  const float lr = slot1 ? (SIDE ? s.lrhs11 : s.lrhs10) : (SIDE ? s.lrhs01 : s.lrhs00);
  const float cur = slot1 ? s.dqd1 : s.dqd0, dinv = slot1 ? s.dinv1 : s.dinv0;
  float dlc = lr - sign * cur * dinv;
  const float sum = fminf(fmaxf(lapp + dlc, 0.0f), LIMIT_MAX_IMPULSE);
  dlc = sum - lapp;
  if (g.lane == 7) {
    lapp = sum;
    if (slot1) s.big1 = fmaxf(s.big1, fabsf(dlc)); else s.big0 = fmaxf(s.big0, fabsf(dlc));
  }
  const float sdl = sign * g.shfl(dlc, 7);
  s.dqd0 += A0[DOF] * sdl;
  s.dqd1 += A1[DOF] * sdl;
}

template <class SM>
__device__ __forceinline__ void limit_rows(const Grp& g, SolverLane& s, const SM& sm, unsigned lact, bool forward, int dof0, const float* A0, const float* A1) {
  if ((lact & ~FINGER_LIMIT_BITS) == 0u) {
    const unsigned f = lact >> (2 * K_F1);
    if (forward) {
      if (f & 1u) finger_limit_row<7, 0>(g, s, A0, A1);
      if (f & 2u) finger_limit_row<7, 1>(g, s, A0, A1);
      if (f & 4u) finger_limit_row<8, 0>(g, s, A0, A1);
      if (f & 8u) finger_limit_row<8, 1>(g, s, A0, A1);
    } else {
      if (f & 8u) finger_limit_row<8, 1>(g, s, A0, A1);
      if (f & 4u) finger_limit_row<8, 0>(g, s, A0, A1);
      if (f & 2u) finger_limit_row<7, 1>(g, s, A0, A1);
      if (f & 1u) finger_limit_row<7, 0>(g, s, A0, A1);
    }
    return;
  }
  unsigned todo = lact;
  while (todo) {
    const int id = forward ? __ffs(todo) - 1 : 31 - __clz(todo);
    todo &= ~(1u << id);
    limit_row(g, s, sm, id, dof0);
  }
}

// ---- finger-table contact rows -------------------------------------------------------------------------
// Row set-up: lane c builds the three rows (normal, two tangents) of cached contact point c on its own --
// the arm axes, M^-1 and the joint velocities are read from shared memory -- so the (up to) 8 points are set
// up in parallel and without a single shuffle.  J of an arm joint at contact point w is
// d . (a_j x (w - p_j)) = a_j . ((w - Pref) x d) + d . vv_j with the published vv_j = (p_j - Pref) x a_j.
__device__ __noinline__ void contact_row_setup(EnvSmem& sm, int c, int n0) {
  const float* hd = sm.hand;
  M3 Rg;
  Rg.r0 = v3(hd[0], hd[1], hd[2]); Rg.r1 = v3(hd[3], hd[4], hd[5]); Rg.r2 = v3(hd[6], hd[7], hd[8]);
  const V3 pf1 = v3(hd[9], hd[10], hd[11]), pf2 = v3(hd[12], hd[13], hd[14]), Pref = v3(hd[15], hd[16], hd[17]);
  const V3 ax1 = v3(hd[18], hd[19], hd[20]), ax2 = -ax1;
  const int k = c < n0 ? 0 : 1, i = k ? c - n0 : c;
  const float* mp = sm.man + k * MAN_WORDS + 1 + 10 * i;
  const V3 lA = v3(mp[0], mp[1], mp[2]), nB = v3(mp[6], mp[7], mp[8]);
  const float dist = mp[9];
  const V3 wr = mul(Rg, lA) + (k ? pf2 : pf1) - Pref;  // contact point on the finger, relative to Pref
  V3 t1, t2;
  plane_space(nB, t1, t2);
#pragma unroll 1
  for (int kk = 0; kk < 3; kk++) {
    const V3 d = kk == 0 ? nB : (kk == 1 ? t1 : t2);
    const V3 m = cross(wr, d);
    float J[ND];  // statically indexed everywhere below: stays in registers
#pragma unroll
    for (int j = 0; j < 7; j++) {
      const float4 p0 = *reinterpret_cast<const float4*>(sm.pub[j]);
      const float2 p1 = *reinterpret_cast<const float2*>(sm.pub[j] + 4);
      J[j] = p0.x * m.x + p0.y * m.y + p0.z * m.z + p0.w * d.x + p1.x * d.y + p1.y * d.z;
    }
    J[7] = k == 0 ? dot(d, ax1) : 0.0f;
    J[8] = k == 1 ? dot(d, ax2) : 0.0f;
    float* row = sm.rows[c * 3 + kk];
    float denom = 0.0f, rel_vel = 0.0f;
#pragma unroll
    for (int j = 0; j < ND; j++) { row[R_J + j] = J[j]; rel_vel += J[j] * sm.vq[j]; }
    // rolled on purpose: this function's register needs leak into the step kernel through the call ABI
    // (unrolling it costs the caller 60 bytes of spills in its hot loop)
#pragma unroll 1
    for (int r = 0; r < ND; r++) {  // M^-1 J^T, one row of M^-1 (three 16-byte loads) per iteration
      const float4* mr = reinterpret_cast<const float4*>(sm.minv + r * MINV_LD);
      const float4 m0 = mr[0], m1 = mr[1], m2 = mr[2];
      const float acc = (m0.x * J[0] + m0.y * J[1] + m0.z * J[2]) + (m0.w * J[3] + m1.x * J[4] + m1.y * J[5]) + (m1.z * J[6] + m1.w * J[7] + m2.x * J[8]);
      row[R_MJ + r] = acc;
      denom += row[R_J + r] * acc;
    }
    const float dinv = 1.0f / denom;
    float rhs;
    if (kk == 0) {
      const float pen = dist + LINEAR_SLOP;
      float pos_err = 0.0f, vel_err = -rel_vel;
      if (pen > 0.0f) vel_err -= pen * INV_DT; else pos_err = -pen * CONTACT_ERP * INV_DT;
      rhs = (pos_err + vel_err) * dinv;
    } else rhs = -rel_vel * dinv;
    row[R_RHS] = rhs; row[R_DINV] = dinv; row[R_DENOM] = denom;
    sm.app[0][c * 3 + kk] = 0.0f;
  }
}

// Row set-up of the one-block tasks: lane c (and again lane c for point c + 8) builds the three rows of cached
// point c.  Pairs in manifold order (pmg_sim.cuh pair_info<1>): finger1-table, finger2-table, table-block,
// floor-block, finger1-block, finger2-block; body A is a finger (robot end: J over the 9 dofs) or a static box
// (no end), body B is the table (no end) or the block (end: d and r_B x d).
// SPILL: the point's records go to the global scratch (c >= SPTS); its own copy of the code, so that the usual
// points keep shared-memory stores and loads.
// S x with S = I_world^-1/2 = R diag(1/sqrt(I)) R^T = k1 + (k3 - k1) a a^T for a body of revolution about its axis a
static_assert(PMG_PUCK_MASS == PMG_BLOCK_MASS, "the puck reuses BLOCK_INV_MASS");
__device__ __forceinline__ V3 puck_scale(V3 x, V3 a) {
  const float pin[3] = PMG_PUCK_INERTIA;
  const float k1 = rsqrtf(pin[0]), k3 = rsqrtf(pin[2]);
  return k1 * x + ((k3 - k1) * dot(a, x)) * a;
}

template <bool SPILL, class SM>
__device__ __noinline__ void contact_row_setup_blk(SM& sm, int c) {
  // which pair / which point of it
  int k = 0, i = c;
#pragma unroll 1
  for (; k < SM::NPAIRS; k++) {
    const int n = __float_as_int(sm.man[k * MAN_WORDS]);
    if (i < n) break;
    i -= n;
  }
  const PairInfo pi = pair_info<SM::NB>(k);
  const bool robotA = pi.ka == G_FINGER1 || pi.ka == G_FINGER2, blockB = pi.kb == G_BLOCK;
  const float* hd = sm.hand;
  M3 Rg;
  Rg.r0 = v3(hd[0], hd[1], hd[2]); Rg.r1 = v3(hd[3], hd[4], hd[5]); Rg.r2 = v3(hd[6], hd[7], hd[8]);
  const V3 pf1 = v3(hd[9], hd[10], hd[11]), pf2 = v3(hd[12], hd[13], hd[14]), Pref = v3(hd[15], hd[16], hd[17]);
  const V3 ax1 = v3(hd[18], hd[19], hd[20]), ax2 = -ax1;
  const float* bk = sm.blk;
  const float* mp = sm.man + k * MAN_WORDS + 1 + 10 * i;
  const V3 lA = v3(mp[0], mp[1], mp[2]), lB = v3(mp[3], mp[4], mp[5]), nB = v3(mp[6], mp[7], mp[8]);
  const float dist = mp[9];
  // contact point on A relative to Pref (robot end) and on B relative to the block centre (block end)
  const V3 wr = robotA ? mul(Rg, lA) + (pi.ka == G_FINGER1 ? pf1 : pf2) - Pref : v3(0, 0, 0);
  M3 Rb;
  Rb.r0 = v3(bk[BK_R], bk[BK_R + 1], bk[BK_R + 2]); Rb.r1 = v3(bk[BK_R + 3], bk[BK_R + 4], bk[BK_R + 5]); Rb.r2 = v3(bk[BK_R + 6], bk[BK_R + 7], bk[BK_R + 8]);
  const V3 rB = blockB ? mul(Rb, lB) : v3(0, 0, 0);
  const V3 bv = v3(bk[BK_V], bk[BK_V + 1], bk[BK_V + 2]), bw = v3(bk[BK_W], bk[BK_W + 1], bk[BK_W + 2]);
  V3 t1, t2;
  plane_space(nB, t1, t2);
  const float mu = geom_friction(pi.ka, SM::PUCK) * geom_friction(pi.kb, SM::PUCK);
  const V3 axb = v3(bk[BK_R + 2], bk[BK_R + 5], bk[BK_R + 8]);  // the puck's axis (third column of its orientation)
#pragma unroll 1
  for (int kk = 0; kk < 3; kk++) {
    const V3 d = kk == 0 ? nB : (kk == 1 ? t1 : t2);
    float* row = SPILL ? sm.spill_row(c * 3 + kk) : sm.rows[SPILL ? 0 : c * 3 + kk];
    const V3 ang0 = cross(rB, d);
    // Puck: rows carry S (r_B x d) with S = I_world^-1/2 and the sweeps solve for S^-1 dw (puck_scale), which keeps
    // the record and the isotropic update of the cubes: (r x d) . dw = (S r x d) . (S^-1 dw), dw += lam I^-1 (r x d).
    const V3 ang = SM::PUCK ? puck_scale(ang0, axb) : ang0;
    float denom = blockB ? BLOCK_INV_MASS + (SM::PUCK ? 1.0f : BLOCK_INV_INERTIA) * dot(ang, ang) : 0.0f;
    float rel_vel = blockB ? -(dot(d, bv) + dot(ang0, bw)) : 0.0f;
    if (robotA) {
      const V3 m = cross(wr, d);
      float J[ND];  // statically indexed everywhere below: stays in registers
#pragma unroll
      for (int j = 0; j < 7; j++) {
        const float4 p0 = *reinterpret_cast<const float4*>(sm.pub[j]);
        const float2 p1 = *reinterpret_cast<const float2*>(sm.pub[j] + 4);
        J[j] = p0.x * m.x + p0.y * m.y + p0.z * m.z + p0.w * d.x + p1.x * d.y + p1.y * d.z;
      }
      J[7] = pi.ka == G_FINGER1 ? dot(d, ax1) : 0.0f;
      J[8] = pi.ka == G_FINGER2 ? dot(d, ax2) : 0.0f;
#pragma unroll
      for (int j = 0; j < ND; j++) { row[R_J + j] = J[j]; rel_vel += J[j] * sm.vq[j]; }
#pragma unroll 1
      for (int r = 0; r < ND; r++) {  // M^-1 J^T, one row of M^-1 (three 16-byte loads) per iteration
        const float4* mr = reinterpret_cast<const float4*>(sm.minv + r * MINV_LD);
        const float4 m0 = mr[0], m1 = mr[1], m2 = mr[2];
        const float acc = (m0.x * J[0] + m0.y * J[1] + m0.z * J[2]) + (m0.w * J[3] + m1.x * J[4] + m1.y * J[5]) + (m1.z * J[6] + m1.w * J[7] + m2.x * J[8]);
        row[R_MJ + r] = acc;
        denom += row[R_J + r] * acc;
      }
    } else {  // static box against the block: no robot end (the sweeps skip this part of the record; zeros for the spilled form)
      const float4 z = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
      for (int j = 0; j < 6; j++) reinterpret_cast<float4*>(row)[j] = z;
    }
    const float dinv = 1.0f / denom;
    float rhs;
    if (kk == 0) {
      const float pen = dist + LINEAR_SLOP;
      float pos_err = 0.0f, vel_err = -rel_vel;
      if (pen > 0.0f) vel_err -= pen * INV_DT; else pos_err = -pen * CONTACT_ERP * INV_DT;
      rhs = (pos_err + vel_err) * dinv;
    } else rhs = -rel_vel * dinv;
    row[R_RHS] = rhs; row[R_DINV] = dinv; row[R_DENOM] = denom; row[R_MU] = mu;
    row[R_BD] = d.x; row[R_BD + 1] = d.y; row[R_BD + 2] = d.z; row[R_BD + 3] = 0.0f;
    row[R_BA] = ang.x; row[R_BA + 1] = ang.y; row[R_BA + 2] = ang.z; row[R_BA + 3] = blockB ? 1.0f : 0.0f;
    sm.app[0][c * 3 + kk] = 0.0f;
  }
}

// One contact row as six 16-byte shared-memory loads: [J0..J8 rhs dinv app] [MJ0..MJ8 denom app2 .]
struct RowVec { float4 a, b, c; };
__device__ __forceinline__ RowVec load3(const float* p) {
  const float4* q = reinterpret_cast<const float4*>(p);
  RowVec r; r.a = q[0]; r.b = q[1]; r.c = q[2];
  return r;
}
__device__ __forceinline__ float row_dot(const RowVec& v, const float* dq) {
  return (v.a.x * dq[0] + v.a.y * dq[1] + v.a.z * dq[2]) + (v.a.w * dq[3] + v.b.x * dq[4] + v.b.y * dq[5]) + (v.b.z * dq[6] + v.b.w * dq[7] + v.c.x * dq[8]);
}
__device__ __forceinline__ void row_axpy(const RowVec& v, float s, float* dq) {
  dq[0] += v.a.x * s; dq[1] += v.a.y * s; dq[2] += v.a.z * s; dq[3] += v.a.w * s; dq[4] += v.b.x * s;
  dq[5] += v.b.y * s; dq[6] += v.b.z * s; dq[7] += v.b.w * s; dq[8] += v.c.x * s;
}

// One Gauss-Seidel pass over the contact rows (all normals, then the friction pairs with the implicit cone).
// Every lane of the octet does the same arithmetic on a replicated delta-velocity vector: no shuffles on
// this path, which is usually executed by one octet of a warp while the others wait.
// Out of line on purpose: it keeps this (rarely executed) code away from the instruction stream of the
// contact-free solver loop.  Delta velocities come in and go out through sm.vq.
// Block end point of a contact row (NBLK > 0): the block is always body B of its pairs, so it sees the
// impulse with the opposite sign; the cubes have isotropic inertia, so M^-1 J^T of the block end is J scaled by
// 1/m (linear) and 1/I (angular) and needs no storage (as in the thread-per-env kernels).
struct BlkVec { float4 d, a; };  // d = (dx dy dz .), a = (r_B x d, has_block)
__device__ __forceinline__ BlkVec load_blk(const float* row) {
  const float4* q = reinterpret_cast<const float4*>(row + R_BD);
  BlkVec r; r.d = q[0]; r.a = q[1];
  return r;
}
__device__ __forceinline__ float blk_dot(const BlkVec& b, const float* dv) {  // dv = block delta (lin[3], ang[3])
  return b.a.w * ((b.d.x * dv[0] + b.d.y * dv[1] + b.d.z * dv[2]) + (b.a.x * dv[3] + b.a.y * dv[4] + b.a.z * dv[5]));
}
template <bool PUCK>  // the puck's rows carry the inertia in their scaled lever arm (contact_row_setup_blk)
__device__ __forceinline__ void blk_axpy(const BlkVec& b, float lam, float* dv) {
  const float lm = lam * b.a.w * BLOCK_INV_MASS, li = lam * b.a.w * (PUCK ? 1.0f : BLOCK_INV_INERTIA);
  dv[0] -= lm * b.d.x; dv[1] -= lm * b.d.y; dv[2] -= lm * b.d.z;
  dv[3] -= li * b.a.x; dv[4] -= li * b.a.y; dv[5] -= li * b.a.z;
}

template <class SM>
__device__ __noinline__ float contact_sweep(Grp g, SM& sm, int nrow_it) {  // nrow | iteration parity << 8 | c1 << 16 | c2 << 24
  constexpr bool BLK = SM::NB > 0;  // rows may have a block end point: the vector gains the block's delta velocities
  const int nrow = nrow_it & 0xff, it = nrow_it >> 8;
  // The accumulated impulses are double buffered (read buffer / write buffer swap every iteration, every lane
  // stores the same value), so no lane waits for another inside the row loops.
  const float* app_rd = sm.app[it & 1];
  float* app_wr = sm.app[(it & 1) ^ 1];
  float dq[ND], dv[6];
#pragma unroll
  for (int j = 0; j < ND; j++) dq[j] = sm.vq[j];
#pragma unroll
  for (int j = 0; j < 6; j++) dv[j] = BLK ? sm.vq[(BLK ? ND : 0) + j] : 0.0f;
  float cres = 0.0f;
  // The loop bodies are lambdas with compile-time flags for the two ends of a row: points come in pair order, so the
  // first c1 have a robot end only (finger-table), those up to c2 a block end only (table-block, floor-block), the
  // rest both (finger-block); a resting block costs 12 instead of 30 multiply-adds per row.  Spilled rows (their own
  // loop: a pointer selected per row would make every row load a generic-address load) take the general form.
  const int c1 = BLK ? (nrow_it >> 16) & 0xff : nrow, c2 = BLK ? (nrow_it >> 24) & 0xff : nrow;
  auto ld4 = [](const float* p) { return *reinterpret_cast<const float4*>(p); };
  auto normal_row = [&](auto robot, auto block, const float* row, int c) {
    constexpr bool RB = decltype(robot)::value, BK = decltype(block)::value;
    const float4 jc = ld4(row + R_J + 8), mc = ld4(row + R_MJ + 8);  // (J8 rhs dinv .) (MJ8 denom mu .)
    RowVec j, mj;
    BlkVec bk;
    float v = 0.0f;
    if (RB) { j.a = ld4(row + R_J); j.b = ld4(row + R_J + 4); j.c = jc; mj.a = ld4(row + R_MJ); mj.b = ld4(row + R_MJ + 4); mj.c = mc; v = row_dot(j, dq); }
    if (BK) { bk = load_blk(row); v -= blk_dot(bk, dv); }
    const float app = app_rd[c * 3];
    float dl = jc.y - v * jc.z;   // rhs - (J . dq) dinv
    const float sum = fminf(fmaxf(app + dl, 0.0f), 1e10f);
    dl = sum - app;
    if (RB) row_axpy(mj, dl, dq);
    if (BK) blk_axpy<SM::PUCK>(bk, dl, dv);
    const float rr = dl * mc.y;                // denom
    cres = fmaxf(cres, rr * rr);
    app_wr[c * 3] = sum;
  };
  auto friction_rows = [&](auto robot, auto block, const float* ra, const float* rb, int c) {  // implicit friction cone: both tangent rows of a point together
    constexpr bool RB = decltype(robot)::value, BK = decltype(block)::value;
    const float total = app_wr[c * 3];
    const float appA = app_rd[c * 3 + 1], appB = app_rd[c * 3 + 2];
    float sA = appA, sB = appB;
    if (total > 0.0f) {
      const float4 jca = ld4(ra + R_J + 8), jcb = ld4(rb + R_J + 8), mca = ld4(ra + R_MJ + 8), mcb = ld4(rb + R_MJ + 8);
      RowVec ja, jb, mja, mjb;
      BlkVec bka, bkb;
      float vA = 0.0f, vB = 0.0f;
      if (RB) {
        ja.a = ld4(ra + R_J); ja.b = ld4(ra + R_J + 4); ja.c = jca; jb.a = ld4(rb + R_J); jb.b = ld4(rb + R_J + 4); jb.c = jcb;
        mja.a = ld4(ra + R_MJ); mja.b = ld4(ra + R_MJ + 4); mja.c = mca; mjb.a = ld4(rb + R_MJ); mjb.b = ld4(rb + R_MJ + 4); mjb.c = mcb;
        vA = row_dot(ja, dq); vB = row_dot(jb, dq);
      }
      if (BK) { bka = load_blk(ra); bkb = load_blk(rb); vA -= blk_dot(bka, dv); vB -= blk_dot(bkb, dv); }
      const float mu = BLK ? mca.z : (float)PMG_FINGER_FRICTION * (float)PMG_TABLE_FRICTION;  // R_MU
      const float lim = mu * total;
      float dA = jca.y - vA * jca.z, dB = jcb.y - vB * jcb.z;
      sA = appA + dA; sB = appB + dB;
      const float s2 = sA * sA + sB * sB;
      if (s2 >= lim * lim) {
        // |lim sin(atan2(sA, sB))| = lim |sA| / sqrt(sA^2 + sB^2), likewise the cosine for sB
        const float sc = s2 > 0.0f ? lim * rsqrtf(s2) : 0.0f;
        const float cA = fabsf(sA) * sc, cB = s2 > 0.0f ? fabsf(sB) * sc : lim;
        sA = fminf(fmaxf(sA, -cA), cA);
        sB = fminf(fmaxf(sB, -cB), cB);
        dA = sA - appA; dB = sB - appB;
      }
      if (RB) { row_axpy(mja, dA, dq); row_axpy(mjb, dB, dq); }
      if (BK) { blk_axpy<SM::PUCK>(bka, dA, dv); blk_axpy<SM::PUCK>(bkb, dB, dv); }
      const float r1_ = dA * mca.y, r2_ = dB * mcb.y;
      cres = fmaxf(cres, fmaxf(r1_ * r1_, r2_ * r2_));
    }
    app_wr[c * 3 + 1] = sA; app_wr[c * 3 + 2] = sB;  // carried over unchanged while the point is open
  };
  using T = std::true_type;
  using F = std::false_type;
  const int nsh = nrow < SM::SPTS ? nrow : SM::SPTS;  // points whose rows are in shared memory
  // one loop per kind (not one loop with a three-way branch): the four environments of a warp then run the same
  // form together even when their point counts differ
  const int e1 = c1 < nsh ? c1 : nsh, e2 = c2 < nsh ? c2 : nsh;
#ifdef PMG_COOP_TIMING
  const long long t_n0 = clock64();
#endif
#pragma unroll 1
  for (int c = 0; c < e1; c++) normal_row(T(), F(), sm.rows[c * 3], c);
  if constexpr (BLK) {
#pragma unroll 1
    for (int c = e1; c < e2; c++) normal_row(F(), T(), sm.rows[c * 3], c);
#pragma unroll 1
    for (int c = e2; c < nsh; c++) normal_row(T(), T(), sm.rows[c * 3], c);
  }
  if constexpr (SM::MAXPTS > SM::SPTS) {
#pragma unroll 1
    for (int c = SM::SPTS; c < nrow; c++) normal_row(T(), T(), sm.spill_row(c * 3), c);
  }
#ifdef PMG_COOP_TIMING
  const long long t_n1 = clock64();
#endif
  g.sync();  // the new normal impulses bound the friction rows
#ifdef PMG_COOP_TIMING
  const long long t_f0 = clock64();
#endif
#pragma unroll 1
  for (int c = 0; c < e1; c++) friction_rows(T(), F(), sm.rows[c * 3 + 1], sm.rows[c * 3 + 2], c);
  if constexpr (BLK) {
#pragma unroll 1
    for (int c = e1; c < e2; c++) friction_rows(F(), T(), sm.rows[c * 3 + 1], sm.rows[c * 3 + 2], c);
#pragma unroll 1
    for (int c = e2; c < nsh; c++) friction_rows(T(), T(), sm.rows[c * 3 + 1], sm.rows[c * 3 + 2], c);
  }
  if constexpr (SM::MAXPTS > SM::SPTS) {
#pragma unroll 1
    for (int c = SM::SPTS; c < nrow; c++) friction_rows(T(), T(), sm.spill_row(c * 3 + 1), sm.spill_row(c * 3 + 2), c);
  }
#ifdef PMG_COOP_TIMING
  if (g.lane == 0) {  // [12] normal loops, [13] friction loops, [14] row visits (normal + friction pair = 2 per point), [15] calls
    const long long t_f1 = clock64();
    atomicAdd(&pmg::g_coop_cycles[12], (unsigned long long)(t_n1 - t_n0)); atomicAdd(&pmg::g_coop_cycles[13], (unsigned long long)(t_f1 - t_f0));
    atomicAdd(&pmg::g_coop_cycles[14], (unsigned long long)(2 * nrow)); atomicAdd(&pmg::g_coop_cycles[15], 1ull);
  }
#endif
  g.sync();  // every lane has read sm.vq
  if (g.lane == 0) {
#pragma unroll
    for (int j = 0; j < ND; j++) sm.vq[j] = dq[j];
    if (BLK) {
#pragma unroll
      for (int j = 0; j < 6; j++) sm.vq[(BLK ? ND : 0) + j] = dv[j];
    }
  }
  g.sync();
  return cres;
}

// ---- multi-block environments (BlockStack / BlockRearrange, NBLK > 1) ------------------------------------------------
// Row set-up: as contact_row_setup_blk, but either end may be a block (block-block pairs have two) and the blocks are
// looked up by index.  Signs as in the thread-per-env kernels (pmg_sim.cuh row_velocity / row_apply): end A sees
// +impulse, end B -impulse.
// Static points (table / floor against block b) get a compact record in sm.srows[b]: their rows touch that block only.
// The other points -- "general": a finger or a second block at the other end -- keep the 40-float record, the first
// SPTS of them in shared memory, the rest in the global spill.
template <int NB>
__device__ __forceinline__ bool static_pair(int k) { return k >= 2 && k < 2 + 4 * NB && ((k - 2) & 3) < 2; }

// Static point `slot` (0..3: the table points, then the floor points) of block b: pair k, point i of its manifold,
// point c of the pair-ordered list (impulse arrays).
template <class SM>
__device__ __noinline__ void contact_row_setup_static(SM& sm, int c, int b, int slot, int k, int i) {
  const float* mp = sm.man + k * MAN_WORDS + 1 + 10 * i;
  const V3 lB = v3(mp[3], mp[4], mp[5]), nB = v3(mp[6], mp[7], mp[8]);
  const float dist = mp[9];
  V3 t1, t2;
  plane_space(nB, t1, t2);
  const float mu = geom_friction((k - 2) & 3 ? G_FLOOR : G_TABLE) * geom_friction(G_BLOCK);
  const float* bk = sm.blk + 24 * b;
  M3 R;
  R.r0 = v3(bk[BK_R], bk[BK_R + 1], bk[BK_R + 2]); R.r1 = v3(bk[BK_R + 3], bk[BK_R + 4], bk[BK_R + 5]); R.r2 = v3(bk[BK_R + 6], bk[BK_R + 7], bk[BK_R + 8]);
  const V3 rB = mul(R, lB);
  const V3 bv = v3(bk[BK_V], bk[BK_V + 1], bk[BK_V + 2]), bw = v3(bk[BK_W], bk[BK_W + 1], bk[BK_W + 2]);
#pragma unroll 1
  for (int kk = 0; kk < 3; kk++) {
    const V3 d = kk == 0 ? nB : (kk == 1 ? t1 : t2);
    float* row = sm.srow((b * 4 + slot) * 3 + kk);
    const V3 xb = cross(rB, d);
    const float denom = BLOCK_INV_MASS + BLOCK_INV_INERTIA * dot(xb, xb);
    const float rel_vel = -(dot(d, bv) + dot(xb, bw));
    const float dinv = 1.0f / denom;
    float rhs;
    if (kk == 0) {
      const float pen = dist + LINEAR_SLOP;
      float pos_err = 0.0f, vel_err = -rel_vel;
      if (pen > 0.0f) vel_err -= pen * INV_DT; else pos_err = -pen * CONTACT_ERP * INV_DT;
      rhs = (pos_err + vel_err) * dinv;
    } else rhs = -rel_vel * dinv;
    row[0] = d.x; row[1] = d.y; row[2] = d.z; row[3] = rhs; row[4] = xb.x; row[5] = xb.y; row[6] = xb.z; row[7] = dinv;
    row[8] = denom; row[9] = mu;
    sm.app[0][c * 3 + kk] = 0.0f;
  }
}

// General point gi (a finger or a second block at the other end): the 40-float record, the first SPTS in shared memory.
template <class SM>
__device__ __noinline__ void contact_row_setup_general(SM& sm, int gidx) {
  const int c = sm.glist[gidx], k = sm.ppair[gidx], i = sm.pidx[gidx];
  const PairInfo pi = pair_info<SM::NB>(k);
  const bool robotA = geom_robot(pi.ka), blockA = pi.ka == G_BLOCK, blockB = pi.kb == G_BLOCK;
  const float* mp = sm.man + k * MAN_WORDS + 1 + 10 * i;
  const V3 lA = v3(mp[0], mp[1], mp[2]), lB = v3(mp[3], mp[4], mp[5]), nB = v3(mp[6], mp[7], mp[8]);
  const float dist = mp[9];
  V3 t1, t2;
  plane_space(nB, t1, t2);
  const float mu = geom_friction(pi.ka) * geom_friction(pi.kb);
  V3 rB = v3(0, 0, 0), bv = v3(0, 0, 0), bw = v3(0, 0, 0);
  if (blockB) {
    const float* bk = sm.blk + 24 * pi.ib;
    M3 R;
    R.r0 = v3(bk[BK_R], bk[BK_R + 1], bk[BK_R + 2]); R.r1 = v3(bk[BK_R + 3], bk[BK_R + 4], bk[BK_R + 5]); R.r2 = v3(bk[BK_R + 6], bk[BK_R + 7], bk[BK_R + 8]);
    rB = mul(R, lB);
    bv = v3(bk[BK_V], bk[BK_V + 1], bk[BK_V + 2]); bw = v3(bk[BK_W], bk[BK_W + 1], bk[BK_W + 2]);
  }
  const float* hd = sm.hand;
  M3 Rg;
  Rg.r0 = v3(hd[0], hd[1], hd[2]); Rg.r1 = v3(hd[3], hd[4], hd[5]); Rg.r2 = v3(hd[6], hd[7], hd[8]);
  const V3 pf1 = v3(hd[9], hd[10], hd[11]), pf2 = v3(hd[12], hd[13], hd[14]), Pref = v3(hd[15], hd[16], hd[17]);
  const V3 ax1 = v3(hd[18], hd[19], hd[20]), ax2 = -ax1;
  // contact point on the robot body relative to Pref (the gripper base's own frame sits AT Pref; its rows have no finger column)
  const V3 wr = robotA ? mul(Rg, lA) + (pi.ka == G_GBASE ? v3(0, 0, 0) : (pi.ka == G_FINGER1 ? pf1 : pf2) - Pref) : v3(0, 0, 0);
  V3 rA = v3(0, 0, 0), av = v3(0, 0, 0), aw = v3(0, 0, 0);
  if (blockA) {
    const float* bk = sm.blk + 24 * pi.ia;
    M3 R;
    R.r0 = v3(bk[BK_R], bk[BK_R + 1], bk[BK_R + 2]); R.r1 = v3(bk[BK_R + 3], bk[BK_R + 4], bk[BK_R + 5]); R.r2 = v3(bk[BK_R + 6], bk[BK_R + 7], bk[BK_R + 8]);
    rA = mul(R, lA);
    av = v3(bk[BK_V], bk[BK_V + 1], bk[BK_V + 2]); aw = v3(bk[BK_W], bk[BK_W + 1], bk[BK_W + 2]);
  }
#pragma unroll 1
  for (int kk = 0; kk < 3; kk++) {
    const V3 d = kk == 0 ? nB : (kk == 1 ? t1 : t2);
    float* row = gidx < SM::SPTS ? sm.rows[gidx * 3 + kk] : sm.spill_row(gidx * 3 + kk);
    const V3 xa = cross(rA, d), xb = cross(rB, d);
    float denom = 0.0f, rel_vel = 0.0f;
    if (blockA) { denom += BLOCK_INV_MASS + BLOCK_INV_INERTIA * dot(xa, xa); rel_vel += dot(d, av) + dot(xa, aw); }
    if (blockB) { denom += BLOCK_INV_MASS + BLOCK_INV_INERTIA * dot(xb, xb); rel_vel -= dot(d, bv) + dot(xb, bw); }
    if (robotA) {
      const V3 m = cross(wr, d);
      float J[ND];
#pragma unroll
      for (int j = 0; j < 7; j++) {
        const float4 p0 = *reinterpret_cast<const float4*>(sm.pub[j]);
        const float2 p1 = *reinterpret_cast<const float2*>(sm.pub[j] + 4);
        J[j] = p0.x * m.x + p0.y * m.y + p0.z * m.z + p0.w * d.x + p1.x * d.y + p1.y * d.z;
      }
      J[7] = pi.ka == G_FINGER1 ? dot(d, ax1) : 0.0f;
      J[8] = pi.ka == G_FINGER2 ? dot(d, ax2) : 0.0f;
#pragma unroll
      for (int j = 0; j < ND; j++) { row[R_J + j] = J[j]; rel_vel += J[j] * sm.vq[j]; }
#pragma unroll 1
      for (int r = 0; r < ND; r++) {
        const float4* mr = reinterpret_cast<const float4*>(sm.minv + r * MINV_LD);
        const float4 m0 = mr[0], m1 = mr[1], m2 = mr[2];
        const float acc = (m0.x * J[0] + m0.y * J[1] + m0.z * J[2]) + (m0.w * J[3] + m1.x * J[4] + m1.y * J[5]) + (m1.z * J[6] + m1.w * J[7] + m2.x * J[8]);
        row[R_MJ + r] = acc;
        denom += row[R_J + r] * acc;
      }
    }  // rows without a robot end leave J / M^-1 J^T unwritten: the sweeps only read them when R_BD + 3 says so
    const float dinv = 1.0f / denom;
    float rhs;
    if (kk == 0) {
      const float pen = dist + LINEAR_SLOP;
      float pos_err = 0.0f, vel_err = -rel_vel;
      if (pen > 0.0f) vel_err -= pen * INV_DT; else pos_err = -pen * CONTACT_ERP * INV_DT;
      rhs = (pos_err + vel_err) * dinv;
    } else rhs = -rel_vel * dinv;
    row[R_RHS] = rhs; row[R_DINV] = dinv; row[R_DENOM] = denom; row[R_MU] = mu;
    row[R_BD] = d.x; row[R_BD + 1] = d.y; row[R_BD + 2] = d.z; row[R_BD + 3] = robotA ? 1.0f : 0.0f;
    row[R_BA] = xb.x; row[R_BA + 1] = xb.y; row[R_BA + 2] = xb.z; row[R_BA + 3] = blockB ? 1.0f : 0.0f;
    row[R_AA] = xa.x; row[R_AA + 1] = xa.y; row[R_AA + 2] = xa.z; row[R_AA + 3] = blockA ? 1.0f : 0.0f;
    row[R_IDX] = __int_as_float(blockA ? pi.ia : -1); row[R_IDX + 1] = __int_as_float(blockB ? pi.ib : -1);
    sm.app[0][c * 3 + kk] = 0.0f;
  }
}

// One Gauss-Seidel pass for multi-block environments.  The joint delta velocities are replicated as in contact_sweep;
// the blocks' are not (a register array cannot be indexed by a run-time block number): LANE b KEEPS BLOCK b's six
// deltas.
//   * STATIC rows (table / floor against block b) involve block b only, so LANE b SOLVES THEM ON ITS OWN, the blocks in
//     parallel: a stack scene at rest is 16 such points = 48 rows per iteration, formerly swept one after the other by
//     the whole octet (76 % of the substep, 430 cycles per row with two run-time-source shuffle groups each).
//     Gauss-Seidel visits rows in Bullet's order -- all normal rows in pair order, then all friction rows -- and two
//     rows commute when they share no velocity; block b's static pairs precede every other pair that touches block b
//     (its finger pairs, the block-block pairs), and they share nothing with the static rows of other blocks or with
//     the finger-table rows, so "static rows of all blocks side by side, then the remaining rows in order" produces the
//     same iterates.
//   * the remaining rows (a finger or a second block at the other end) are visited by the whole octet in pair order: a
//     row fetches its one or two blocks with run-time-source shuffles (legal on the diverged path: the whole octet
//     takes it) and the owner lanes apply the impulse.
// nrow_it: iteration parity << 8 | number of general points << 16; srange: this lane's static points [s0, s1) as
// s0 | s1 << 8 (a contiguous run of the pair-ordered point list: table-block, then floor-block), both from the substep.
template <class SM>
__device__ __noinline__ float contact_sweep_multi(Grp g, SM& sm, int nrow_it, int srange) {
  constexpr int NB = SM::NB;
  const int it = (nrow_it >> 8) & 1, ngen = (nrow_it >> 16) & 0xff;
  const int lane = g.lane;
  const float* app_rd = sm.app[it & 1];
  float* app_wr = sm.app[(it & 1) ^ 1];
  float dq[ND], bd[6];
  if (ngen) {  // only rows with a robot end read the joint delta velocities (uniform over the octet)
#pragma unroll
    for (int j = 0; j < ND; j++) dq[j] = sm.vq[j];
  }
#pragma unroll
  for (int j = 0; j < 6; j++) bd[j] = lane < NB ? sm.vq[ND + 6 * lane + j] : 0.0f;
  float cres = 0.0f;
  auto ld4 = [](const float* p) { return *reinterpret_cast<const float4*>(p); };
  const int s0 = srange & 0xff, s1 = srange >> 8;
  float sn[4];  // normal impulses of this lane's static points after the normal pass
  const float* srow = sm.srow((lane < NB ? lane : 0) * 12);  // this lane's compact static rows: [slot * 3 + kk][12]
  // ---- normal rows ----
  // static: the block is end B (it sees -impulse).  At most 4 points per block (contact_row_setup_static): the records
  // are fetched up front -- they do not depend on the velocities -- so that only the arithmetic sits on the dependent
  // chain from one row to the next.
  {
    float4 D[4], X[4];
    float den[4], ap[4];
#pragma unroll
    for (int i = 0; i < 4; i++)
      if (s0 + i < s1) {
        const float* row = srow + i * 3 * SM::SROW_W;
        D[i] = ld4(row); X[i] = ld4(row + 4); den[i] = row[8]; ap[i] = app_rd[(s0 + i) * 3];  // D.w = rhs, X.w = 1 / denominator
      }
#pragma unroll
    for (int i = 0; i < 4; i++)
      if (s0 + i < s1) {
        const float4 d = D[i], xb = X[i];
        const float v = -((d.x * bd[0] + d.y * bd[1] + d.z * bd[2]) + (xb.x * bd[3] + xb.y * bd[4] + xb.z * bd[5]));
        float dl = d.w - v * xb.w;
        const float sum = fminf(fmaxf(ap[i] + dl, 0.0f), 1e10f);
        dl = sum - ap[i];
        const float lm = dl * BLOCK_INV_MASS, li = dl * BLOCK_INV_INERTIA;
        bd[0] -= lm * d.x; bd[1] -= lm * d.y; bd[2] -= lm * d.z; bd[3] -= li * xb.x; bd[4] -= li * xb.y; bd[5] -= li * xb.z;
        const float rr = dl * den[i];
        cres = fmaxf(cres, rr * rr);
        app_wr[(s0 + i) * 3] = sum;
        ap[i] = sum;  // the normal impulse the friction rows of this point are bounded by
      }
    // the friction records of the static points, fetched while the general normal rows run
#pragma unroll
    for (int i = 0; i < 4; i++) { sn[i] = s0 + i < s1 ? ap[i] : 0.0f; }
  }
  // velocity of a general row's block ends: + end A, - end B
  auto block_vel = [&](const float4& d, const float4& xa, const float4& xb, int ia, int ib) {
    float v = 0.0f;
    if (ia >= 0) {
      const float l0 = g.shfl(bd[0], ia), l1 = g.shfl(bd[1], ia), l2 = g.shfl(bd[2], ia), w0 = g.shfl(bd[3], ia), w1 = g.shfl(bd[4], ia), w2 = g.shfl(bd[5], ia);
      v += (d.x * l0 + d.y * l1 + d.z * l2) + (xa.x * w0 + xa.y * w1 + xa.z * w2);
    }
    if (ib >= 0) {
      const float l0 = g.shfl(bd[0], ib), l1 = g.shfl(bd[1], ib), l2 = g.shfl(bd[2], ib), w0 = g.shfl(bd[3], ib), w1 = g.shfl(bd[4], ib), w2 = g.shfl(bd[5], ib);
      v -= (d.x * l0 + d.y * l1 + d.z * l2) + (xb.x * w0 + xb.y * w1 + xb.z * w2);
    }
    return v;
  };
  auto block_apply = [&](const float4& d, const float4& xa, const float4& xb, int ia, int ib, float dl) {
    const float lm = dl * BLOCK_INV_MASS, li = dl * BLOCK_INV_INERTIA;
    if (lane == ia) { bd[0] += lm * d.x; bd[1] += lm * d.y; bd[2] += lm * d.z; bd[3] += li * xa.x; bd[4] += li * xa.y; bd[5] += li * xa.z; }
    if (lane == ib) { bd[0] -= lm * d.x; bd[1] -= lm * d.y; bd[2] -= lm * d.z; bd[3] -= li * xb.x; bd[4] -= li * xb.y; bd[5] -= li * xb.z; }
  };
  auto normal_row = [&](const float* row, int c) {
    const float4 jc = ld4(row + R_J + 8), mc = ld4(row + R_MJ + 8), d = ld4(row + R_BD), xb = ld4(row + R_BA), xa = ld4(row + R_AA);
    const int ia = __float_as_int(row[R_IDX]), ib = __float_as_int(row[R_IDX + 1]);
    const bool rob = d.w != 0.0f;
    RowVec j, mj;
    float v = 0.0f;
    if (rob) { j.a = ld4(row + R_J); j.b = ld4(row + R_J + 4); j.c = jc; mj.a = ld4(row + R_MJ); mj.b = ld4(row + R_MJ + 4); mj.c = mc; v = row_dot(j, dq); }
    v += block_vel(d, xa, xb, ia, ib);
    const float app = app_rd[c * 3];
    float dl = jc.y - v * jc.z;
    const float sum = fminf(fmaxf(app + dl, 0.0f), 1e10f);
    dl = sum - app;
    if (rob) row_axpy(mj, dl, dq);
    block_apply(d, xa, xb, ia, ib, dl);
    const float rr = dl * mc.y;
    cres = fmaxf(cres, rr * rr);
    app_wr[c * 3] = sum;
  };
  auto friction_rows = [&](const float* ra, const float* rb, int c) {
    const float total = app_wr[c * 3];
    const float appA = app_rd[c * 3 + 1], appB = app_rd[c * 3 + 2];
    float sA = appA, sB = appB;
    if (total > 0.0f) {  // uniform over the octet: every lane holds the same impulses
      const float4 jca = ld4(ra + R_J + 8), jcb = ld4(rb + R_J + 8), mca = ld4(ra + R_MJ + 8), mcb = ld4(rb + R_MJ + 8);
      const float4 da = ld4(ra + R_BD), xba = ld4(ra + R_BA), xaa = ld4(ra + R_AA), db = ld4(rb + R_BD), xbb = ld4(rb + R_BA), xab = ld4(rb + R_AA);
      const int ia = __float_as_int(ra[R_IDX]), ib = __float_as_int(ra[R_IDX + 1]);
      const bool rob = da.w != 0.0f;
      RowVec ja, jb, mja, mjb;
      float vA = 0.0f, vB = 0.0f;
      if (rob) {
        ja.a = ld4(ra + R_J); ja.b = ld4(ra + R_J + 4); ja.c = jca; jb.a = ld4(rb + R_J); jb.b = ld4(rb + R_J + 4); jb.c = jcb;
        mja.a = ld4(ra + R_MJ); mja.b = ld4(ra + R_MJ + 4); mja.c = mca; mjb.a = ld4(rb + R_MJ); mjb.b = ld4(rb + R_MJ + 4); mjb.c = mcb;
        vA = row_dot(ja, dq); vB = row_dot(jb, dq);
      }
      vA += block_vel(da, xaa, xba, ia, ib); vB += block_vel(db, xab, xbb, ia, ib);
      const float lim = mca.z * total;  // R_MU
      float dA = jca.y - vA * jca.z, dB = jcb.y - vB * jcb.z;
      sA = appA + dA; sB = appB + dB;
      const float s2 = sA * sA + sB * sB;
      if (s2 >= lim * lim) {
        const float sc = s2 > 0.0f ? lim * rsqrtf(s2) : 0.0f;
        const float cA = fabsf(sA) * sc, cB = s2 > 0.0f ? fabsf(sB) * sc : lim;
        sA = fminf(fmaxf(sA, -cA), cA);
        sB = fminf(fmaxf(sB, -cB), cB);
        dA = sA - appA; dB = sB - appB;
      }
      if (rob) { row_axpy(mja, dA, dq); row_axpy(mjb, dB, dq); }
      block_apply(da, xaa, xba, ia, ib, dA); block_apply(db, xab, xbb, ia, ib, dB);
      const float r1_ = dA * mca.y, r2_ = dB * mcb.y;
      cres = fmaxf(cres, fmaxf(r1_ * r1_, r2_ * r2_));
    }
    app_wr[c * 3 + 1] = sA; app_wr[c * 3 + 2] = sB;
  };
  // the other rows, in pair order, by the whole octet (the pair counts are uniform over the octet)
  {
    const int nsh = ngen < SM::SPTS ? ngen : SM::SPTS;
#pragma unroll 1
    for (int gi = 0; gi < nsh; gi++) normal_row(sm.rows[gi * 3], sm.glist[gi]);  // glist: point index (impulse arrays)
#pragma unroll 1
    for (int gi = SM::SPTS; gi < ngen; gi++) normal_row(sm.spill_row(gi * 3), sm.glist[gi]);
  }
  // A static point's impulses are private to its lane; a general point's normal impulse was just stored by every lane
  // of the octet (the same value) and is read back by the friction pass: one octet barrier orders those accesses
  // (racecheck reports the unordered identical stores as hazards), only when there are such points.
  if (ngen) g.sync();
  // ---- friction rows: the two tangent rows of a point projected together onto the cone ----
  {
    // (The records are fetched point by point: holding all four points' friction records at once -- 84 registers --
    // pushed this function past what the call site leaves free, and every call then saved and restored ~40 registers
    // through local memory, whose L1 share next to 200 KB of shared memory is tiny: 24 % of the block_stack step.)
#pragma unroll
    for (int i = 0; i < 4; i++)
      if (s0 + i < s1) {
        const float* ra = srow + (i * 3 + 1) * SM::SROW_W;
        const float* rb = ra + SM::SROW_W;
        const float4 da = ld4(ra), xba = ld4(ra + 4), db = ld4(rb), xbb = ld4(rb + 4);
        const float2 ma = *reinterpret_cast<const float2*>(ra + 8);  // denominator, friction coefficient
        const float denb = rb[8];
        const float apa = app_rd[(s0 + i) * 3 + 1], apb = app_rd[(s0 + i) * 3 + 2];
        const float total = sn[i];
        float sA = apa, sB = apb;
        if (total > 0.0f) {
          const float vA = -((da.x * bd[0] + da.y * bd[1] + da.z * bd[2]) + (xba.x * bd[3] + xba.y * bd[4] + xba.z * bd[5]));
          const float vB = -((db.x * bd[0] + db.y * bd[1] + db.z * bd[2]) + (xbb.x * bd[3] + xbb.y * bd[4] + xbb.z * bd[5]));
          const float lim = ma.y * total;
          float dA = da.w - vA * xba.w, dB = db.w - vB * xbb.w;
          sA = apa + dA; sB = apb + dB;
          const float s2 = sA * sA + sB * sB;
          if (s2 >= lim * lim) {
            const float sc = s2 > 0.0f ? lim * rsqrtf(s2) : 0.0f;
            const float cA = fabsf(sA) * sc, cB = s2 > 0.0f ? fabsf(sB) * sc : lim;
            sA = fminf(fmaxf(sA, -cA), cA);
            sB = fminf(fmaxf(sB, -cB), cB);
            dA = sA - apa; dB = sB - apb;
          }
          const float lmA = dA * BLOCK_INV_MASS, liA = dA * BLOCK_INV_INERTIA, lmB = dB * BLOCK_INV_MASS, liB = dB * BLOCK_INV_INERTIA;
          bd[0] -= lmA * da.x; bd[1] -= lmA * da.y; bd[2] -= lmA * da.z; bd[3] -= liA * xba.x; bd[4] -= liA * xba.y; bd[5] -= liA * xba.z;
          bd[0] -= lmB * db.x; bd[1] -= lmB * db.y; bd[2] -= lmB * db.z; bd[3] -= liB * xbb.x; bd[4] -= liB * xbb.y; bd[5] -= liB * xbb.z;
          const float r1_ = dA * ma.x, r2_ = dB * denb;
          cres = fmaxf(cres, fmaxf(r1_ * r1_, r2_ * r2_));
        }
        app_wr[(s0 + i) * 3 + 1] = sA; app_wr[(s0 + i) * 3 + 2] = sB;
      }
  }
  {
    const int nsh = ngen < SM::SPTS ? ngen : SM::SPTS;
#pragma unroll 1
    for (int gi = 0; gi < nsh; gi++) friction_rows(sm.rows[gi * 3 + 1], sm.rows[gi * 3 + 2], sm.glist[gi]);
#pragma unroll 1
    for (int gi = SM::SPTS; gi < ngen; gi++) friction_rows(sm.spill_row(gi * 3 + 1), sm.spill_row(gi * 3 + 2), sm.glist[gi]);
  }
  if (lane < NB) {  // private to the lane: read back by the same lane in the next iteration / at the integration
#pragma unroll
    for (int j = 0; j < 6; j++) sm.vq[ND + 6 * lane + j] = bd[j];
  }
  if (ngen) {
    g.sync();  // every lane has read sm.vq
    if (lane == 0) {
#pragma unroll
      for (int j = 0; j < ND; j++) sm.vq[j] = dq[j];
    }
    g.sync();
  }
  return cres;
}

// Development aid (tools/coop_timing.py builds a separate library with -DPMG_COOP_TIMING): cycles spent by
// octets with cached contacts in [0] the whole substep, [1] narrowphase, [2] row set-up, [3] contact sweeps,
// and [4] the number of such substeps; [5] whole substep / [6] count for contact-free octets.
#ifdef PMG_COOP_TIMING
#define PMG_T(var) const long long var = clock64()
#define PMG_TADD(slot, cyc) do { if (g.lane == 0) atomicAdd(&pmg::g_coop_cycles[slot], (unsigned long long)(cyc)); } while (0)
#else
#define PMG_T(var)
#define PMG_TADD(slot, cyc)
#endif

// ---- one 2 ms substep ---------------------------------------------------------------------------------
template <class SM>
__device__ void substep(const Grp& g, SM& sm, Lane& L) {
  constexpr bool BLK = SM::NB > 0;
  constexpr bool MULTI = SM::NB > 1;  // BlockStack / BlockRearrange: lane b owns block b
  PMG_T(t_begin);
#ifdef PMG_COOP_TIMING
  long long t_sweeps = 0;
#endif
  const int lane = g.lane;
  const bool arm = lane < 7, hand = lane == 7;
  if (BLK && (MULTI ? lane < SM::NB : lane == 0)) {  // the block's orientation matrix for the narrowphase and the rows (visible after the sync below)
    float* bk = sm.blk + (MULTI ? 24 * lane : 0);
    const M3 Rb = quat_to_m3(bk[BK_QUAT], bk[BK_QUAT + 1], bk[BK_QUAT + 2], bk[BK_QUAT + 3]);
    bk[BK_R] = Rb.r0.x; bk[BK_R + 1] = Rb.r0.y; bk[BK_R + 2] = Rb.r0.z; bk[BK_R + 3] = Rb.r1.x; bk[BK_R + 4] = Rb.r1.y;
    bk[BK_R + 5] = Rb.r1.z; bk[BK_R + 6] = Rb.r2.x; bk[BK_R + 7] = Rb.r2.y; bk[BK_R + 8] = Rb.r2.z;
  }
  // 1. link frames (scan), joint axes
  M3 R; V3 p;
  chain_fk(g, L, arm ? L.q0 : 0.0f, R, p);
  const V3 a = arm ? col(R, 2) : v3(0, 0, 0);
  const V3 Pref = g.shfl(p, 7);  // gripper-base origin: reference point of every moment below
  // 2. velocities and zero-qdd accelerations as prefix sums
  const V3 aq = L.qd0 * a;
  const V3 w = g.scan(aq), wp = w - aq;
  const V3 tal = cross(wp, aq);
  const V3 al = g.scan(tal), alp = al - tal;
  V3 pprev = g.up(p, 1);
  if (lane == 0) pprev = v3(0, 0, 0);
  const V3 r = p - pprev;
  const V3 wxr = cross(wp, r);
  const V3 vo = g.scan(wxr);
  V3 ao = g.scan(cross(alp, r) + cross(wp, wxr));
  ao.z += GRAVITY;
  // 3. Newton-Euler of the lane's own body about its COM (Bullet's link damping as an external force)
  V3 Fs, Ns, hs;
  Sym3 Is;
  {
    const float mass = L.lc[LC_MASS];
    const V3 rc = mul(R, lc3(L.lc, LC_COM));
    const Sym3 Iw = world_inertia(R, lc3(L.lc, LC_INERTIA));
    const V3 wxrc = cross(w, rc);
    const V3 a_c = ao + cross(al, rc) + cross(w, wxrc);
    const V3 v_c = vo + wxrc;
    const float kl = LINK_DAMPING + LINK_DAMPING * norm(v_c), ka = LINK_DAMPING + LINK_DAMPING * norm(w);
    const V3 Fc = mass * a_c + (mass * kl) * v_c;
    const V3 Iww = mul(Iw, w);
    const V3 Nc = mul(Iw, al) + cross(w, Iww) + ka * Iww;
    const V3 c = (p - Pref) + rc;
    Fs = Fc; Ns = Nc + cross(c, Fc); hs = mass * c;
    Is = sym_add(Iw, point_inertia(mass, c));
  }
  // 3b. the two fingers (same orientation as the gripper base, zero COM offset), by every lane
  const M3 Rg = g.shfl(R, 7);
  const V3 wg = g.shfl(w, 7), alg = g.shfl(al, 7), aog = g.shfl(ao, 7), vog = g.shfl(vo, 7);
  const float qf1 = g.shfl(L.q0, 7), qf2 = g.shfl(L.q1, 7), qdf1 = g.shfl(L.qd0, 7), qdf2 = g.shfl(L.qd1, 7);
  const V3 ay = col(Rg, 1);
  const V3 ax1 = -ay, ax2 = ay;  // prismatic axes (0,-1,0) / (0,+1,0) of the gripper base
  const float mf = c_mass[PMG_BODY_FINGER1];
  V3 r1, r2, F1, F2, nf1, lf1, nf2, lf2;
  {
    r1 = mul(Rg, v3(c_jxyz[PMG_BODY_FINGER1][0], c_jxyz[PMG_BODY_FINGER1][1], c_jxyz[PMG_BODY_FINGER1][2])) + qf1 * ax1;
    r2 = mul(Rg, v3(c_jxyz[PMG_BODY_FINGER2][0], c_jxyz[PMG_BODY_FINGER2][1], c_jxyz[PMG_BODY_FINGER2][2])) + qf2 * ax2;
    const V3 aq1 = qdf1 * ax1, aq2 = qdf2 * ax2;
    const V3 wxr1 = cross(wg, r1), wxr2 = cross(wg, r2);
    const V3 a1 = aog + cross(alg, r1) + cross(wg, wxr1) + 2.0f * cross(wg, aq1);
    const V3 a2 = aog + cross(alg, r2) + cross(wg, wxr2) + 2.0f * cross(wg, aq2);
    const V3 v1 = vog + wxr1 + aq1, v2 = vog + wxr2 + aq2;
    const float kl1 = LINK_DAMPING + LINK_DAMPING * norm(v1), kl2 = LINK_DAMPING + LINK_DAMPING * norm(v2);
    F1 = mf * a1 + (mf * kl1) * v1;
    F2 = mf * a2 + (mf * kl2) * v2;
    const Sym3 Iwf = world_inertia(Rg, v3(c_inertia[PMG_BODY_FINGER1][0], c_inertia[PMG_BODY_FINGER1][1], c_inertia[PMG_BODY_FINGER1][2]));
    const float ka = LINK_DAMPING + LINK_DAMPING * norm(wg);
    const V3 Iww = mul(Iwf, wg);
    const V3 Ncf = mul(Iwf, alg) + cross(wg, Iww) + ka * Iww;
    // momentum of a unit finger velocity about Pref: l = m ax, n = m r x ax
    lf1 = mf * ax1; nf1 = mf * cross(r1, ax1);
    lf2 = mf * ax2; nf2 = mf * cross(r2, ax2);
    if (hand) {  // the gripper base carries its fingers
      Fs += F1 + F2;
      Ns += Ncf + Ncf + cross(r1, F1) + cross(r2, F2);
      hs += mf * (r1 + r2);
      Is = sym_add(Is, sym_add(sym_add(Iwf, Iwf), sym_add(point_inertia(mf, r1), point_inertia(mf, r2))));
    }
  }
  const V3 pf1 = Pref + r1, pf2 = Pref + r2;
  if (lane == 2) {  // stash the gripper frame for the contact-row set-up (lanes 0 and 1 are about to collide)
    float* hd = sm.hand;
    hd[0] = Rg.r0.x; hd[1] = Rg.r0.y; hd[2] = Rg.r0.z; hd[3] = Rg.r1.x; hd[4] = Rg.r1.y; hd[5] = Rg.r1.z;
    hd[6] = Rg.r2.x; hd[7] = Rg.r2.y; hd[8] = Rg.r2.z;
    hd[9] = pf1.x; hd[10] = pf1.y; hd[11] = pf1.z; hd[12] = pf2.x; hd[13] = pf2.y; hd[14] = pf2.z;
    hd[15] = Pref.x; hd[16] = Pref.y; hd[17] = Pref.z; hd[18] = ax1.x; hd[19] = ax1.y; hd[20] = ax1.z;
  }
  // collision detection of the two finger-table pairs: lane k runs pair k on the shared-memory manifold
  PMG_T(t_col0);
  if (BLK) g.sync();  // block pose of this substep (integration of the last one, orientation matrix above)
  if constexpr (MULTI) {
    // lane k runs pairs k, k + 8, ...: the same geometry look-up as the thread-per-env kernels, blocks from shared memory
    ManRef mr; mr.man = sm.man; mr.stride = 1;
    constexpr int SCR_STRIDE = SM::SCRATCH_FLOATS / SM::NSCR / 4 * 4;
    static_assert(SCR_STRIDE * sizeof(float) >= sizeof(BoxScratch), "one narrowphase scratch area per lane");
    BoxScratch& scr = *reinterpret_cast<BoxScratch*>(&sm.rows[0][0] + lane * SCR_STRIDE);
    const float tc[3] = PMG_TABLE_CENTER, fc[3] = PMG_FLOOR_CENTER;
    // Position p of the schedule -> pair: the table-block pairs (always touching in a scene at rest: a full SAT +
    // clipping each) come first, one per lane, then finger-table, finger-block, block-block, floor-block, base-block.
    auto scheduled_pair = [](int p) {
      constexpr int NB = SM::NB, NBB = NB * (NB - 1) / 2;
      if (p < NB) return 2 + 4 * p;
      p -= NB;
      if (p < 2) return p;
      p -= 2;
      if (p < 2 * NB) return 2 + 4 * (p >> 1) + 2 + (p & 1);
      p -= 2 * NB;
      if (p < NBB) return 2 + 4 * NB + p;
      p -= NBB;
      if (p < NB) return 2 + 4 * p + 1;
      return 2 + 4 * NB + NBB + (p - NB);  // gripper base - block
    };
    // The gripper-base pairs (the last NB positions) all fail the broadphase while the cylinder's world box -- grown by
    // the largest box a cube can fill and the margins -- stays above every block: the usual case (a block on the table
    // is 3 cm below the base at the lowest tip height).  One comparison per block then replaces their round of the loop.
    int npos = SM::NPAIRS;
    {
      const float ez = (fabsf(Rg.r2.x) + fabsf(Rg.r2.y)) * (float)PMG_GBASE_RADIUS + fabsf(Rg.r2.z) * (float)PMG_GBASE_HALF_LEN;
      const float low = Pref.z - (ez + 1.7320508f * BLOCK_HALF + 2 * BROADPHASE_MARGIN);
      bool reach = false;
#pragma unroll
      for (int b = 0; b < SM::NB; b++) reach = reach || !(sm.blk[24 * b + BK_POS + 2] < low);
      if (!reach) {
        npos -= SM::NB;
        if (lane < SM::NB) {  // what collide_pair's own broadphase would do: forget the cached points
          float& cnt = sm.man[(2 + 4 * SM::NB + SM::NB * (SM::NB - 1) / 2 + lane) * MAN_WORDS];
          if (__float_as_int(cnt) != 0) cnt = __int_as_float(0);
        }
      }
    }
#pragma unroll 1
    for (int pos = lane; pos < npos; pos += GL) {
      const int k = scheduled_pair(pos);
      const PairInfo pi = pair_info<SM::NB>(k);
      V3 pa, pb;
      M3 Ra = m3_identity(), Rb = m3_identity();
      if (pi.ka == G_FINGER1 || pi.ka == G_FINGER2) { pa = pi.ka == G_FINGER1 ? pf1 : pf2; Ra = Rg; }
      else if (pi.ka == G_GBASE) { pa = Pref; Ra = Rg; }  // the cylinder is centred on the gripper-base frame
      else if (pi.ka == G_TABLE) pa = v3(tc[0], tc[1], tc[2]);
      else if (pi.ka == G_FLOOR) pa = v3(fc[0], fc[1], fc[2]);
      else {
        const float* bk = sm.blk + 24 * pi.ia;
        pa = v3(bk[BK_POS], bk[BK_POS + 1], bk[BK_POS + 2]);
        Ra.r0 = v3(bk[BK_R], bk[BK_R + 1], bk[BK_R + 2]); Ra.r1 = v3(bk[BK_R + 3], bk[BK_R + 4], bk[BK_R + 5]); Ra.r2 = v3(bk[BK_R + 6], bk[BK_R + 7], bk[BK_R + 8]);
      }
      if (pi.kb == G_TABLE) pb = v3(tc[0], tc[1], tc[2]);
      else {
        const float* bk = sm.blk + 24 * pi.ib;
        pb = v3(bk[BK_POS], bk[BK_POS + 1], bk[BK_POS + 2]);
        Rb.r0 = v3(bk[BK_R], bk[BK_R + 1], bk[BK_R + 2]); Rb.r1 = v3(bk[BK_R + 3], bk[BK_R + 4], bk[BK_R + 5]); Rb.r2 = v3(bk[BK_R + 6], bk[BK_R + 7], bk[BK_R + 8]);
      }
      collide_pair(mr, k, pa, Ra, geom_half(pi.ka), geom_anchor(pi.ka), geom_static(pi.ka), pb, Rb, geom_half(pi.kb), geom_anchor(pi.kb), scr, false, geom_static(pi.kb),
                   pi.ka == G_GBASE);
    }
  } else if (lane < SM::NPAIRS) {
    ManRef mr; mr.man = sm.man; mr.stride = 1;
    const float tc[3] = PMG_TABLE_CENTER, th[3] = PMG_TABLE_HALF, fh[3] = PMG_FINGER_HALF;
    // the narrowphase work arrays live in the (not yet used) contact-row area of shared memory
    constexpr int SCR_STRIDE = SM::SPTS * 3 * SM::ROW_W / SM::NSCR / 4 * 4;
    BoxScratch& scr = *reinterpret_cast<BoxScratch*>(&sm.rows[0][0] + lane * SCR_STRIDE);
    if constexpr (!BLK) {
      collide_pair(mr, lane, lane == 0 ? pf1 : pf2, Rg, v3(fh[0], fh[1], fh[2]), v3(0, 0, 0), false,
                   v3(tc[0], tc[1], tc[2]), m3_identity(), v3(th[0], th[1], th[2]), geom_anchor(G_TABLE), scr, false, true);
    } else {
      // lane k runs pair k: finger1-table, finger2-table, table-block, floor-block, finger1-block, finger2-block
      const PairInfo pi = pair_info<SM::NB>(lane);
      const float fc[3] = PMG_FLOOR_CENTER;
      const float* bk = sm.blk;
      const bool fingerA = pi.ka == G_FINGER1 || pi.ka == G_FINGER2;
      const V3 tcv = table_center(SM::PUCK);  // Slide: the long table
      V3 pa = fingerA ? (pi.ka == G_FINGER1 ? pf1 : pf2) : (pi.ka == G_TABLE ? tcv : v3(fc[0], fc[1], fc[2]));
      M3 Ra = fingerA ? Rg : m3_identity();
      V3 pb = pi.kb == G_TABLE ? tcv : v3(bk[BK_POS], bk[BK_POS + 1], bk[BK_POS + 2]);
      M3 Rb = m3_identity();
      if (pi.kb == G_BLOCK) {
        Rb.r0 = v3(bk[BK_R], bk[BK_R + 1], bk[BK_R + 2]); Rb.r1 = v3(bk[BK_R + 3], bk[BK_R + 4], bk[BK_R + 5]); Rb.r2 = v3(bk[BK_R + 6], bk[BK_R + 7], bk[BK_R + 8]);
      }
      collide_pair(mr, lane, pa, Ra, geom_half(pi.ka, SM::PUCK), geom_anchor(pi.ka), geom_static(pi.ka), pb, Rb, geom_half(pi.kb, SM::PUCK), geom_anchor(pi.kb), scr,
                   SM::PUCK && pi.kb == G_BLOCK, geom_static(pi.kb));
    }
  }
  PMG_T(t_col1);
  // 4. subtree wrenches and composite inertias: suffix sums over the chain
  Fs = g.rscan(Fs); Ns = g.rscan(Ns); hs = g.rscan(hs);
  Is.xx = g.rscan(Is.xx); Is.xy = g.rscan(Is.xy); Is.xz = g.rscan(Is.xz);
  Is.yy = g.rscan(Is.yy); Is.yz = g.rscan(Is.yz); Is.zz = g.rscan(Is.zz);
  // 5. joint-space inertia matrix rows and bias forces
  const V3 vv = cross(p - Pref, a);  // velocity of the reference point per unit joint rate
  const V3 n_own = mul(Is, a) + cross(hs, vv);
  const V3 l_own = L.lc[LC_MSUB] * vv + cross(a, hs);
  const float b0 = arm ? dot(a, Ns) + dot(vv, Fs) : dot(ax1, F1);
  const float b1 = dot(ax2, F2);
  if (arm) {
    float* pb = sm.pub[lane];
    pb[0] = a.x; pb[1] = a.y; pb[2] = a.z; pb[3] = vv.x; pb[4] = vv.y; pb[5] = vv.z;
  }
  g.sync();
  // Lane i evaluates M[i][j] = s_j . (n_i, l_i) for the joints j <= i it sits below (its subtree carries
  // their motion) and writes both triangle positions into the shared-memory matrix; the finger rows are
  // the finger columns of the arm rows.  Then every lane reads its full row(s) back.
  float* Ms = sm.minv;
#pragma unroll
  for (int j = 0; j < 7; j++) {
    const float* pb = sm.pub[j];
    const float val = pb[0] * n_own.x + pb[1] * n_own.y + pb[2] * n_own.z + pb[3] * l_own.x + pb[4] * l_own.y + pb[5] * l_own.z;
    if (arm && j <= lane) { Ms[lane * MINV_LD + j] = val; Ms[j * MINV_LD + lane] = val; }
  }
  if (arm) {
    const float m7 = dot(a, nf1) + dot(vv, lf1), m8 = dot(a, nf2) + dot(vv, lf2);
    Ms[lane * MINV_LD + 7] = m7; Ms[7 * MINV_LD + lane] = m7;
    Ms[lane * MINV_LD + 8] = m8; Ms[8 * MINV_LD + lane] = m8;
  } else {  // the fingers are siblings: no coupling
    Ms[7 * MINV_LD + 7] = mf; Ms[7 * MINV_LD + 8] = 0.0f; Ms[8 * MINV_LD + 7] = 0.0f; Ms[8 * MINV_LD + 8] = mf;
  }
  g.sync();
  float A0[ND], A1[ND];
#pragma unroll
  for (int j = 0; j < ND; j++) {
    A0[j] = Ms[L.dof0 * MINV_LD + j];
    A1[j] = hand ? Ms[8 * MINV_LD + j] : (j == 8 ? 1.0f : 0.0f);  // arm lanes carry an inert second row
  }
  // 6. M^-1 by Gauss-Jordan sweeps, row k broadcast from its owner (lane 7 owns rows 7 and 8).  The
  // owner's own row equals the broadcast row, so "row -= f * pivot_row" with f = 1 - 1/pivot scales it.
#pragma unroll
  for (int k = 0; k < ND; k++) {
    const int src = k < 8 ? k : 7;
    float rk[ND];
#pragma unroll
    for (int j = 0; j < ND; j++) rk[j] = g.shfl(k == 8 ? A1[j] : A0[j], src);
    const float pinv = 1.0f / rk[k];
    const bool own0 = lane == src && k != 8, own1 = lane == src && k == 8;
    const float f0 = own0 ? 1.0f - pinv : A0[k] * pinv, f1 = own1 ? 1.0f - pinv : A1[k] * pinv;
#pragma unroll
    for (int j = 0; j < ND; j++) {
      if (j == k) continue;
      A0[j] -= f0 * rk[j];
      A1[j] -= f1 * rk[j];
    }
    A0[k] = own0 ? pinv : -f0; A1[k] = own1 ? pinv : -f1;
  }
#pragma unroll
  for (int j = 0; j < ND; j++) {
    sm.minv[L.dof0 * MINV_LD + j] = A0[j];
    if (hand) sm.minv[8 * MINV_LD + j] = A1[j];
  }
  // 7. unconstrained velocity update: qd += dt * M^-1 (tau_damping - bias)
  {
    const float t0 = L.dtau0 - b0, t1 = L.dtau1 - b1;
    float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
    for (int j = 0; j < ND; j++) {
      const float tj = j < 8 ? g.shfl(t0, j) : g.shfl(t1, 7);
      s0 += A0[j] * tj; s1 += A1[j] * tj;
    }
    L.qd0 = fminf(fmaxf(L.qd0 + s0 * DT, -MAX_COORD_VEL), MAX_COORD_VEL);
    L.qd1 = hand ? fminf(fmaxf(L.qd1 + s1 * DT, -MAX_COORD_VEL), MAX_COORD_VEL) : 0.0f;
  }
  sm.vq[L.dof0] = L.qd0;
  if (hand) sm.vq[8] = L.qd1;
  if (BLK && (MULTI ? lane < SM::NB : lane == 0)) {  // free cube: gravity + Bullet's velocity damping; isotropic inertia => no gyro term
    float* bk = sm.blk + (MULTI ? 24 * lane : 0);
    V3 bv = v3(bk[BK_V], bk[BK_V + 1], bk[BK_V + 2]), bw = v3(bk[BK_W], bk[BK_W + 1], bk[BK_W + 2]);
    if constexpr (SM::PUCK) {
      // anisotropic inertia: the gyroscopic term alpha = -I^-1 (w x I w), I = i1 + (i3 - i1) a a^T about the puck's axis a
      // (w x I w = (i3 - i1)(a.w) w x a; I^-1 y = y / i1 + (1/i3 - 1/i1)(a.y) a and a . (w x a) = 0)
      const float pin[3] = PMG_PUCK_INERTIA;
      const V3 axb = v3(bk[BK_R + 2], bk[BK_R + 5], bk[BK_R + 8]);
      const V3 g = ((pin[2] - pin[0]) * dot(axb, bw)) * cross(bw, axb);
      bw -= (DT / pin[0]) * g;
    }
    const float kl = LINK_DAMPING + LINK_DAMPING * norm(bv), ka = LINK_DAMPING + LINK_DAMPING * norm(bw);
    bv += DT * (v3(0, 0, -GRAVITY) - kl * bv);
    bw -= (DT * ka) * bw;
    bk[BK_V] = bv.x; bk[BK_V + 1] = bv.y; bk[BK_V + 2] = bv.z; bk[BK_W] = bw.x; bk[BK_W + 1] = bw.y; bk[BK_W + 2] = bw.z;
#pragma unroll
    for (int j = 0; j < 6; j++) sm.vq[(BLK ? ND : 0) + (MULTI ? 6 * lane : 0) + j] = 0.0f;  // the block's PGS delta velocities
  }
  g.sync();  // minv, vq, the block velocity and the manifolds are visible to the whole octet
  // 8. constraint rows
  SolverLane s;
  unsigned lact = 0;
  {
    s.mdd0 = sm.minv[L.dof0 * (MINV_LD + 1)]; s.mdd1 = A1[8];
    s.dinv0 = 1.0f / s.mdd0; s.dinv1 = 1.0f / s.mdd1;
    // btMultiBodyJointMotor in POSITION_CONTROL (kuka.py:282-301)
    const float tv0 = MOTOR_KP * (L.mt0 - L.q0) * INV_DT + L.qd0 + MOTOR_KD * (0.0f - L.qd0);
    const float tv1 = MOTOR_KP * (L.mt1 - L.q1) * INV_DT + L.qd1 + MOTOR_KD * (0.0f - L.qd1);
    s.rhs0 = (tv0 - L.qd0) * s.dinv0; s.rhs1 = (tv1 - L.qd1) * s.dinv1;
    s.lim0 = L.mi0; s.lim1 = L.mi1; s.app0 = s.app1 = 0.0f;
    // btMultiBodyJointLimitConstraint rows, only when violated
    const float p00 = L.q0 - L.lc[LC_LOWER], p01 = L.lc[LC_UPPER] - L.q0, p10 = L.q1 - c_dof_lower[8], p11 = c_dof_upper[8] - L.q1;
    const bool v00 = !(p00 > 0.0f), v01 = !(p01 > 0.0f), v10 = hand && !(p10 > 0.0f), v11 = hand && !(p11 > 0.0f);
    // the four violation flags of every lane, gathered with one warp reduction (a partial-mask vote per flag
    // was the most expensive instruction of the contact-free path): lane l owns bits 4l .. 4l+3
    const unsigned mine = ((v00 ? 1u : 0u) | (v01 ? 2u : 0u) | (v10 ? 4u : 0u) | (v11 ? 8u : 0u)) << (4 * lane);
    const unsigned viol = g.reduce_or(mine);
    s.lrhs00 = s.lrhs01 = s.lrhs10 = s.lrhs11 = 0.0f;
    if (viol) {
      s.lrhs00 = ((p00 > SPLIT_IMPULSE_PEN_THRESHOLD ? -p00 * CONTACT_ERP * INV_DT : 0.0f) - L.qd0) * s.dinv0;
      s.lrhs01 = ((p01 > SPLIT_IMPULSE_PEN_THRESHOLD ? -p01 * CONTACT_ERP * INV_DT : 0.0f) + L.qd0) * s.dinv0;
      s.lrhs10 = ((p10 > SPLIT_IMPULSE_PEN_THRESHOLD ? -p10 * CONTACT_ERP * INV_DT : 0.0f) - L.qd1) * s.dinv1;
      s.lrhs11 = ((p11 > SPLIT_IMPULSE_PEN_THRESHOLD ? -p11 * CONTACT_ERP * INV_DT : 0.0f) + L.qd1) * s.dinv1;
#pragma unroll
      for (int k = 0; k < ND; k++) {
        const int d = nc_dof(k);
        const unsigned lo = d < 8 ? (viol >> (4 * d)) & 1u : (viol >> 30) & 1u, hi = d < 8 ? (viol >> (4 * d + 1)) & 1u : (viol >> 31) & 1u;
        lact |= (lo << (2 * k)) | (hi << (2 * k + 1));
      }
    }
    s.lapp00 = s.lapp01 = s.lapp10 = s.lapp11 = 0.0f;
    s.dqd0 = s.dqd1 = 0.0f;
  }
  // contact rows: one normal + two tangents per cached manifold point, point c set up by lane c
  const int n0 = __float_as_int(sm.man[0]);
  int nrow = n0 + __float_as_int(sm.man[MAN_WORDS]);  // number of cached contact points (3 rows each)
  int row_kinds = 0;                                  // (first point with a block end) << 16 | (first with both ends) << 24
  int ngen = 0, srange = 0;                           // multi-block scenes: general points, this lane's static points
  int st0[SM::NB > 1 ? SM::NB : 1], ntab[SM::NB > 1 ? SM::NB : 1], nst[SM::NB > 1 ? SM::NB : 1];  // per block: first static point, table points, static points kept
  if (BLK) {
    row_kinds = nrow << 16;
#pragma unroll
    for (int k = 2; k < SM::NPAIRS; k++) {
      if (k == 4) row_kinds |= nrow << 24;
      nrow += __float_as_int(sm.man[k * MAN_WORDS]);
    }
    if (nrow > SM::MAXPTS) {  // points beyond the pool are dropped and counted (pmg_overflow_count)
      if (lane == 0) sm.blk[23] += (float)(nrow - SM::MAXPTS);
      nrow = SM::MAXPTS;
    }
    if constexpr (MULTI) {
      // Point classes, once per substep, in registers: per block the run of its STATIC points (table-block, then
      // floor-block; at most 4 slots -- points beyond them, i.e. table and floor at once, are dropped and counted
      // like the pool overflow) and the number of the other ("general") points.
      constexpr int NBK = SM::NB;
      int c = 0;
#pragma unroll
      for (int b = 0; b < NBK; b++) { st0[b] = 0; ntab[b] = 0; nst[b] = 0; }
#pragma unroll
      for (int k = 0; k < SM::NPAIRS; k++) {
        const int n = __float_as_int(sm.man[k * MAN_WORDS]);
        int ne = nrow - c;  // points beyond the pool were dropped
        ne = ne < 0 ? 0 : (ne < n ? ne : n);
        if (static_pair<NBK>(k)) {
          const int b = (k - 2) >> 2;
          if (((k - 2) & 3) == 0) { st0[b] = c; ntab[b] = ne; nst[b] = ne; }
          else {
            if (lane == 0 && nst[b] + ne > 4) sm.blk[23] += (float)(nst[b] + ne - 4);
            nst[b] = nst[b] + ne < 4 ? nst[b] + ne : 4;
          }
        } else ngen += ne;
        c += n;
      }
      int s0 = 0, s1 = 0;
#pragma unroll
      for (int b = 0; b < NBK; b++) if (lane == b) { s0 = st0[b]; s1 = st0[b] + nst[b]; }
      srange = s0 | (s1 << 8);
      if (ngen) {  // rare in a scene at rest: a finger on a block or on the table, stacked blocks
        if (lane == 0) {
          int c2 = 0, g2 = 0;
#pragma unroll 1
          for (int k = 0; k < SM::NPAIRS; k++) {
            const int n = __float_as_int(sm.man[k * MAN_WORDS]);
            if (!static_pair<NBK>(k))
              for (int i = 0; i < n && c2 + i < nrow; i++) {
                sm.glist[g2] = (unsigned char)(c2 + i); sm.ppair[g2] = (unsigned char)k; sm.pidx[g2] = (unsigned char)i;
                g2++;
              }
            c2 += n;
          }
        }
        g.sync();  // the general-point tables are read by the row set-up of every lane
      }
    }
  }
  PMG_T(t_set0);
  if (nrow) {
    if constexpr (!BLK) {
      if (lane < nrow) contact_row_setup(sm, lane, n0);
    } else {
      if constexpr (MULTI) {
        // static slot s = 4 b + slot over the lanes (a scene at rest: 16 slots, two per lane), then the general points
        for (int sidx = lane; sidx < SM::NB * 4; sidx += GL) {
          const int b = sidx >> 2, slot = sidx & 3;
          int b0 = 0, nt = 0, ns = 0;
#pragma unroll
          for (int bb = 0; bb < SM::NB; bb++) if (bb == b) { b0 = st0[bb]; nt = ntab[bb]; ns = nst[bb]; }
          if (slot < ns) contact_row_setup_static(sm, b0 + slot, b, slot, slot < nt ? 2 + 4 * b : 3 + 4 * b, slot < nt ? slot : slot - nt);
        }
        for (int gi = lane; gi < ngen; gi += GL) contact_row_setup_general(sm, gi);
      } else {
        for (int c = lane; c < nrow; c += GL) {
          if (c < SM::SPTS) contact_row_setup_blk<false>(sm, c);
          else contact_row_setup_blk<true>(sm, c);
        }
      }
    }
    g.sync();
  }
  PMG_T(t_set1);
  // projected Gauss-Seidel: <= 5 iterations, early exit on the largest squared velocity change
  for (int it = 0; it < SOLVER_ITERS; it++) {
    s.big0 = s.big1 = 0.0f;
    if (it & 1) {  // forwards on odd iterations, backwards on even (Bullet's interleaving)
      motor_row<0>(g, s, A0, A1); motor_row<1>(g, s, A0, A1); motor_row<2>(g, s, A0, A1);
      motor_row<3>(g, s, A0, A1); motor_row<4>(g, s, A0, A1); motor_row<5>(g, s, A0, A1);
      motor_row<6>(g, s, A0, A1); motor_row<7>(g, s, A0, A1); motor_row<8>(g, s, A0, A1);
      if (lact) limit_rows(g, s, sm, lact, true, L.dof0, A0, A1);
    } else {
      if (lact) limit_rows(g, s, sm, lact, false, L.dof0, A0, A1);
      motor_row<8>(g, s, A0, A1); motor_row<7>(g, s, A0, A1); motor_row<6>(g, s, A0, A1);
      motor_row<5>(g, s, A0, A1); motor_row<4>(g, s, A0, A1); motor_row<3>(g, s, A0, A1);
      motor_row<2>(g, s, A0, A1); motor_row<1>(g, s, A0, A1); motor_row<0>(g, s, A0, A1);
    }
    const float e0 = s.big0 * s.mdd0, e1 = s.big1 * s.mdd1;
    float res = fmaxf(e0 * e0, e1 * e1);
    if (nrow) {
      // gather the delta velocities, run the contact rows replicated, take the own components back (multi-block scenes
      // whose points are all static leave the joints alone: no exchange, no barriers)
      const bool joints = !MULTI || ngen != 0;
      if (joints) {
        sm.vq[L.dof0] = s.dqd0;
        if (hand) sm.vq[8] = s.dqd1;
        g.sync();
      }
      PMG_T(t_sw0);
      if constexpr (MULTI) res = fmaxf(res, contact_sweep_multi(g, sm, ((it & 1) << 8) | (ngen << 16), srange));
      else res = fmaxf(res, contact_sweep(g, sm, nrow | ((it & 1) << 8) | row_kinds));
#ifdef PMG_COOP_TIMING
      t_sweeps += clock64() - t_sw0;
#endif
      if (joints) {
        s.dqd0 = sm.vq[L.dof0];
        s.dqd1 = sm.vq[8];
      }
    }
    res = g.maxv(res);
    if (res <= RESIDUAL_THRESHOLD) break;
  }
  // 9. semi-implicit Euler
  L.qd0 = fminf(fmaxf(L.qd0 + s.dqd0, -MAX_COORD_VEL), MAX_COORD_VEL);
  L.qd1 = hand ? fminf(fmaxf(L.qd1 + s.dqd1, -MAX_COORD_VEL), MAX_COORD_VEL) : 0.0f;
  L.q0 += L.qd0 * DT;
  L.q1 += L.qd1 * DT;
  if (BLK && (MULTI ? lane < SM::NB : lane == 0)) {  // block: add the solver's delta velocities, integrate (exponential map for the orientation)
    float* bk = sm.blk + (MULTI ? 24 * lane : 0);
    const float* dv = sm.vq + (BLK ? ND : 0) + (MULTI ? 6 * lane : 0);
    V3 bv = v3(bk[BK_V] + dv[0], bk[BK_V + 1] + dv[1], bk[BK_V + 2] + dv[2]);
    V3 dw = v3(dv[3], dv[4], dv[5]);
    if constexpr (SM::PUCK) dw = puck_scale(dw, v3(bk[BK_R + 2], bk[BK_R + 5], bk[BK_R + 8]));  // the sweeps solved for S^-1 dw
    V3 bw = v3(bk[BK_W] + dw.x, bk[BK_W + 1] + dw.y, bk[BK_W + 2] + dw.z);
    bk[BK_V] = bv.x; bk[BK_V + 1] = bv.y; bk[BK_V + 2] = bv.z; bk[BK_W] = bw.x; bk[BK_W + 1] = bw.y; bk[BK_W + 2] = bw.z;
    bk[BK_POS] += DT * bv.x; bk[BK_POS + 1] += DT * bv.y; bk[BK_POS + 2] += DT * bv.z;
    float w = norm(bw);
    if (w * DT > 0.25f * PI_F) w = 0.25f * PI_F / DT;  // ANGULAR_MOTION_THRESHOLD
    const float sn = w < 0.001f ? 0.5f * DT - DT * DT * DT * 0.020833333333f * w * w : sinf(0.5f * w * DT) / w;
    const float ax = bw.x * sn, ay = bw.y * sn, az = bw.z * sn, aw = cosf(0.5f * w * DT);
    const float bx = bk[BK_QUAT], by = bk[BK_QUAT + 1], bz = bk[BK_QUAT + 2], bw_ = bk[BK_QUAT + 3];
    const float nx = aw * bx + ax * bw_ + ay * bz - az * by;
    const float ny = aw * by + ay * bw_ + az * bx - ax * bz;
    const float nz = aw * bz + az * bw_ + ax * by - ay * bx;
    const float nw = aw * bw_ - ax * bx - ay * by - az * bz;
    const float inv = 1.0f / sqrtf(nx * nx + ny * ny + nz * nz + nw * nw);
    bk[BK_QUAT] = nx * inv; bk[BK_QUAT + 1] = ny * inv; bk[BK_QUAT + 2] = nz * inv; bk[BK_QUAT + 3] = nw * inv;
  }
#ifdef PMG_COOP_TIMING
  if (nrow) {
    PMG_TADD(0, clock64() - t_begin); PMG_TADD(1, t_col1 - t_col0); PMG_TADD(2, t_set1 - t_set0); PMG_TADD(3, t_sweeps); PMG_TADD(4, 1);
  } else { PMG_TADD(5, clock64() - t_begin); PMG_TADD(6, 1); PMG_TADD(7, t_col1 - t_col0); }
#endif
}

// ---- persistent manifolds between HBM and shared memory ------------------------------------------------------
// Only what exists is moved: every pair's point count, and the 10 words of each cached point (a contact-free Reach
// environment: 2 words instead of 82; a block-stack scene at rest: 28 counts + 16 points instead of 1148 words).  Words
// beyond a pair's count are never read (manifold_add / manifold_refresh / the row set-up all stop at the count).
template <int NPAIRS>
__device__ __forceinline__ void load_manifolds(const Grp& g, float* sman, const float* mg, size_t B) {
  for (int k = g.lane; k < NPAIRS; k += GL) sman[k * MAN_WORDS] = mg[(size_t)(k * MAN_WORDS) * B];
  g.sync();
#pragma unroll 1
  for (int k = 0; k < NPAIRS; k++) {
    const int n = __float_as_int(sman[k * MAN_WORDS]);
    for (int w = 1 + g.lane; w < 1 + 10 * n; w += GL) sman[k * MAN_WORDS + w] = mg[(size_t)(k * MAN_WORDS + w) * B];
  }
}
template <int NPAIRS>
__device__ __forceinline__ void store_manifolds(const Grp& g, const float* sman, float* mg, size_t B) {
  for (int k = g.lane; k < NPAIRS; k += GL) mg[(size_t)(k * MAN_WORDS) * B] = sman[k * MAN_WORDS];
#pragma unroll 1
  for (int k = 0; k < NPAIRS; k++) {
    const int n = __float_as_int(sman[k * MAN_WORDS]);
    for (int w = 1 + g.lane; w < 1 + 10 * n; w += GL) mg[(size_t)(k * MAN_WORDS + w) * B] = sman[k * MAN_WORDS + w];
  }
}

// ---- fused gather: the octet copies its finished row, reward and flags to the peers' gather buffers -------------
// (multi-GPU sharding, include/pmg.h pmg_step_gather).  The row was just written to this rank's own buffer by lanes
// of this octet; after the octet barrier the 8 lanes copy it with 32-byte segments to every peer over NVLink, then
// lane 0 signs the environment off (gather_arrive).  Nothing happens when no gather is attached (g_n == 0).
__device__ __forceinline__ void gather_push(const Grp& g, const StepIO& io, int env) {
  if (io.g_n == 0 || !io.g_in_step) return;
  g.sync();
  const int W = io.row_width;
  const float* row = io.obs + (size_t)env * W;
  for (int d = 0; d < io.g_n; d++) {
    float* dst = io.g_obs[d] + (size_t)env * W;
    for (int k = g.lane; k < W; k += GL) dst[k] = row[k];
    if (g.lane == 0) { io.g_reward[d][env] = io.reward[env]; io.g_done[d][env] = io.done[env]; io.g_success[d][env] = io.success[env]; }
  }
  __threadfence_system();
  g.sync();
  if (g.lane == 0) gather_arrive(io);
}

// ---- one env.step() of a Reach environment (TASK 0, no blocks) -------------------------------------
// `env` is the environment index, state / manifold are the same [word][env] arrays the thread-per-env
// kernels use, so reset_kernel, pmg_get_state / pmg_set_state and the reward path are shared.
// `lane_consts`: the block's constant table (8 rows of LC_W floats, filled with fill_lane_constants).
// JC: joint control (a compile-time variant, so that the default kernel keeps its register allocation)
template <bool JC>
__device__ void step_env_reach(const Grp& g, EnvSmem& sm, const float* lane_consts, const StepIO& io, int env) {
  using D = Dims<0, 0>;
  const int lane = g.lane;
  const bool arm = lane < 7, hand = lane == 7;
  const size_t B = io.tile;  // distance between consecutive words of this environment (StepIO::tile)
  float* s = io.state + state_off(io, env);
  Lane L;
  L.lc = lane_consts + lane * LC_W;
  L.dof0 = lane;  // lane 7 -> dof 7 (finger1); it also owns dof 8 in slot 1
  L.q0 = s[(ST_Q + lane) * B]; L.qd0 = s[(ST_QD + lane) * B]; L.mt0 = s[(ST_MT + lane) * B]; L.mi0 = s[(ST_MI + lane) * B];
  L.q1 = hand ? s[(ST_Q + 8) * B] : 0.0f; L.qd1 = hand ? s[(ST_QD + 8) * B] : 0.0f;
  L.mt1 = hand ? s[(ST_MT + 8) * B] : 0.0f; L.mi1 = hand ? s[(ST_MI + 8) * B] : 0.0f;
  L.dtau0 = L.dtau1 = 0.0f;
  load_manifolds<EnvSmem::NPAIRS>(g, sm.man, io.manifold + man_off(io, env), B);
  // ---- Kuka.apply_action (kuka.py:167-222) ----
  const float lo[3] = {-0.67f, -0.20f, 0.175f}, hi[3] = {-0.37f, 0.20f, 0.55f};  // kuka.py:40-41
  float ee[3];
  if (JC) {
    // kuka.py:204-206: joint_state_target (kept in the motor targets) += 0.05 a[:7]; no clipping, no IK
#pragma unroll
    for (int k = 0; k < 3; k++) ee[k] = s[(ST_EE + k) * B];
    if (arm) { L.mt0 += io.action[(size_t)env * io.adim + lane] * 0.05f; L.mi0 = ARM_FORCE * OUTER_DT; }
  } else {
#pragma unroll
    for (int k = 0; k < 3; k++) ee[k] = fminf(fmaxf(s[(ST_EE + k) * B] + io.action[(size_t)env * D::A + k] * 0.01f, lo[k]), hi[k]);
    const float tq[4] = {0.f, -1.f, 0.f, 0.f};  // kuka.py:42
    const float qik = inverse_kinematics(g, L, arm ? L.q0 : 0.0f, v3(ee[0], ee[1], ee[2]), tq);
    if (arm) { L.mt0 = qik; L.mi0 = ARM_FORCE * OUTER_DT; }
  }
  g.sync();
  // ---- 5 x stepSimulation (kuka.py:223-225), each 20 substeps of 2 ms ----
  for (int call = 0; call < CALLS_PER_ENV_STEP; call++) {
    L.dtau0 = -L.lc[LC_DAMP] * L.qd0;  // joint damping torque, sampled once per stepSimulation call
    L.dtau1 = 0.0f;
    for (int sub = 0; sub < SUBSTEPS_PER_CALL; sub++) { g.block_sync(); substep(g, sm, L); }
  }
  // ---- observation, reward, flags (kuka_single_step_base_env.py:193-244) ----
  M3 R; V3 p;
  chain_fk(g, L, arm ? L.q0 : 0.0f, R, p);
  s[(ST_Q + lane) * B] = L.q0; s[(ST_QD + lane) * B] = L.qd0; s[(ST_MT + lane) * B] = L.mt0; s[(ST_MI + lane) * B] = L.mi0;
  if (hand) { s[(ST_Q + 8) * B] = L.q1; s[(ST_QD + 8) * B] = L.qd1; s[(ST_MT + 8) * B] = L.mt1; s[(ST_MI + 8) * B] = L.mi1; }
  g.sync();
  store_manifolds<EnvSmem::NPAIRS>(g, sm.man, io.manifold + man_off(io, env), B);
  // joint control prepends the 7 joint positions to observation and policy_state (kuka_single_step_base_env.py:214-216)
  constexpr int jo = JC ? 7 : 0;
  if (JC && arm) {
    float* row = io.obs + (size_t)env * io.row_width;
    row[lane] = L.q0; row[3 + jo + lane] = L.q0;
  }
  if (lane == PMG_BODY_LINK7) {
    const float t[3] = PMG_TIP_OFFSET;
    const V3 tip = p + mul(R, v3(t[0], t[1], t[2]));
    float* row = io.obs + (size_t)env * (D::W + 2 * jo);
    const float* goal = s + (size_t)(ST_BLK) * B;
    const float g0 = goal[0], g1 = goal[B], g2 = goal[2 * B];
    float* ob = row + jo; float* pol = row + 3 + 2 * jo; float* ag = row + 6 + 2 * jo;
    ob[0] = pol[0] = ag[0] = tip.x; ob[1] = pol[1] = ag[1] = tip.y; ob[2] = pol[2] = ag[2] = tip.z;
    ag[3] = g0; ag[4] = g1; ag[5] = g2;
    const float dx = tip.x - g0, dy = tip.y - g1, dz = tip.z - g2;
    const float dist = sqrtf(dx * dx + dy * dy + dz * dz);
#pragma unroll
    for (int k = 0; k < 3; k++) s[(ST_EE + k) * B] = ee[k];
    float* el = s + (size_t)(D::STATE - 1) * B;
    const int elapsed = (int)(*el) + 1;
    *el = (float)elapsed;
    const bool na = dist > io.thr;
    io.reward[env] = io.binary ? -(na ? 1.0f : 0.0f) : -dist;
    io.success[env] = na ? 0 : 1;
    io.done[env] = elapsed >= io.max_steps ? 1 : 0;
  }
  gather_push(g, io, env);
}

// ---- one env.step() of a one-block environment: Push (TASK 1) / PickAndPlace (TASK 2) / Slide (TASK 5) -----
// Same structure as step_env_reach; the block's state and its four extra manifolds live in shared memory for
// the whole step.  Slide (kuka_single_step_envs.py:49-59) is Push on the long table with the puck: the same
// action map, observation and reward, EnvSmemT<1, true> selects its geometry and inertia.
template <int TASK>
__device__ void step_env_block(const Grp& g, EnvSmemT<1, TASK == 5>& sm, const float* lane_consts, const StepIO& io, int env) {
  using SM = EnvSmemT<1, TASK == 5>;
  using D = Dims<TASK == 5 ? 1 : TASK, 1>;
  const int lane = g.lane;
  const bool arm = lane < 7, hand = lane == 7;
  const size_t B = io.tile;  // distance between consecutive words of this environment (StepIO::tile)
  float* s = io.state + state_off(io, env);
  Lane L;
  L.lc = lane_consts + lane * LC_W;
  L.dof0 = lane;
  L.q0 = s[(ST_Q + lane) * B]; L.qd0 = s[(ST_QD + lane) * B]; L.mt0 = s[(ST_MT + lane) * B]; L.mi0 = s[(ST_MI + lane) * B];
  L.q1 = hand ? s[(ST_Q + 8) * B] : 0.0f; L.qd1 = hand ? s[(ST_QD + 8) * B] : 0.0f;
  L.mt1 = hand ? s[(ST_MT + 8) * B] : 0.0f; L.mi1 = hand ? s[(ST_MI + 8) * B] : 0.0f;
  L.dtau0 = L.dtau1 = 0.0f;
  load_manifolds<SM::NPAIRS>(g, sm.man, io.manifold + man_off(io, env), B);
  for (int w = lane; w < 13; w += GL) sm.blk[w] = s[(size_t)(ST_BLK + w) * B];
  if (lane == 0) {
    sm.blk[23] = 0.0f;  // contact points dropped because the row pool was full
    sm.spill = io.row_spill + (size_t)env * SM::SPILL_WORDS;
  }
  // ---- Kuka.apply_action (kuka.py:167-222) ----
  const float* act = io.action + (size_t)env * (io.jc ? io.adim : D::A);
  if (TASK == 2 && hand) {  // grasping: the last action column drives both jaws (kuka.py:169-172)
    const float grip = (act[io.jc ? 7 : 3] + 1.0f) * (GRIPPER_ABS_LIMIT / 2);
    L.mt0 = L.mt1 = grip; L.mi0 = L.mi1 = FINGER_FORCE * OUTER_DT;
  }
  const float lo[3] = {-0.67f, -0.20f, 0.175f}, hi[3] = {-0.37f, 0.20f, 0.55f};  // kuka.py:40-41
  float ee[3];
  if (io.jc) {
    // kuka.py:204-206: joint_state_target (kept in the motor targets) += 0.05 a[:7]; no clipping, no IK
#pragma unroll
    for (int k = 0; k < 3; k++) ee[k] = s[(ST_EE + k) * B];
    if (arm) { L.mt0 += act[lane] * 0.05f; L.mi0 = ARM_FORCE * OUTER_DT; }
  } else {
#pragma unroll
    for (int k = 0; k < 3; k++) ee[k] = fminf(fmaxf(s[(ST_EE + k) * B] + act[k] * 0.01f, lo[k]), hi[k]);
    const float tq[4] = {0.f, -1.f, 0.f, 0.f};  // kuka.py:42
    const float qik = inverse_kinematics(g, L, arm ? L.q0 : 0.0f, v3(ee[0], ee[1], ee[2]), tq);
    if (arm) { L.mt0 = qik; L.mi0 = ARM_FORCE * OUTER_DT; }
  }
  g.sync();
  // ---- 5 x stepSimulation (kuka.py:223-225), each 20 substeps of 2 ms ----
  for (int call = 0; call < CALLS_PER_ENV_STEP; call++) {
    L.dtau0 = -L.lc[LC_DAMP] * L.qd0;  // joint damping torque, sampled once per stepSimulation call
    L.dtau1 = 0.0f;
    for (int sub = 0; sub < SUBSTEPS_PER_CALL; sub++) { g.block_sync(); substep(g, sm, L); }
  }
  // ---- observation, reward, flags (kuka.py:227-256, kuka_single_step_base_env.py:193-244) ----
  M3 R; V3 p;
  chain_fk(g, L, arm ? L.q0 : 0.0f, R, p);
  const V3 a = arm ? col(R, 2) : v3(0, 0, 0);
  const V3 aq = L.qd0 * a;
  const V3 w = g.scan(aq), wp = w - aq;
  V3 pprev = g.up(p, 1);
  if (lane == 0) pprev = v3(0, 0, 0);
  const V3 vo = g.scan(cross(wp, p - pprev));
  s[(ST_Q + lane) * B] = L.q0; s[(ST_QD + lane) * B] = L.qd0; s[(ST_MT + lane) * B] = L.mt0; s[(ST_MI + lane) * B] = L.mi0;
  if (hand) { s[(ST_Q + 8) * B] = L.q1; s[(ST_QD + 8) * B] = L.qd1; s[(ST_MT + 8) * B] = L.mt1; s[(ST_MI + 8) * B] = L.mi1; }
  g.sync();  // the block state of the last substep
  store_manifolds<SM::NPAIRS>(g, sm.man, io.manifold + man_off(io, env), B);
  for (int wd = lane; wd < 13; wd += GL) s[(size_t)(ST_BLK + wd) * B] = sm.blk[wd];
  if (lane == 0 && sm.blk[23] > 0.0f && io.overflow) atomicAdd(io.overflow, (int)sm.blk[23]);
  // tip pose / velocity from lane 6 (link_7); jaw quantities on lane 7 (gripper base), which writes the row
  const float t[3] = PMG_TIP_OFFSET;
  const V3 tip_own = p + mul(R, v3(t[0], t[1], t[2]));
  const V3 tip = g.shfl(tip_own, PMG_BODY_LINK7);
  const V3 tv = g.shfl(vo + cross(w, tip_own - p), PMG_BODY_LINK7), tw = g.shfl(w, PMG_BODY_LINK7);
  // joint control prepends the 7 joint positions to observation and policy_state (kuka_single_step_base_env.py:214-216)
  const int jo = io.jc ? 7 : 0;
  if (io.jc && arm) {
    float* row = io.obs + (size_t)env * io.row_width;
    row[lane] = L.q0; row[D::O + jo + lane] = L.q0;
  }
  if (hand) {
    float closeness = 0.0f, finger_vel = 0.0f;
    if (TASK == 2) {
      const float t1[3] = PMG_TAB1_OFFSET, t2[3] = PMG_TAB2_OFFSET;
      const V3 ay = col(R, 1);  // R = gripper-base frame on this lane; fingers slide along -/+ its y axis
      const V3 j1 = v3(c_jxyz[PMG_BODY_FINGER1][0], c_jxyz[PMG_BODY_FINGER1][1], c_jxyz[PMG_BODY_FINGER1][2]);
      const V3 j2 = v3(c_jxyz[PMG_BODY_FINGER2][0], c_jxyz[PMG_BODY_FINGER2][1], c_jxyz[PMG_BODY_FINGER2][2]);
      const V3 tab1 = mul(R, j1 + v3(t1[0], t1[1], t1[2])) - L.q0 * ay;   // relative to the gripper-base origin
      const V3 tab2 = mul(R, j2 + v3(t2[0], t2[1], t2[2])) + L.q1 * ay;
      closeness = norm(tab1 - tab2);
      // finger_vel = (v_base - v_tab1).y with v_tab1 = v_base + w x (tab1 - p_base) + qd_finger1 * axis1 (kuka.py:240-242)
      const V3 rel = cross(w, tab1) - L.qd0 * ay;
      finger_vel = -rel.y;
    }
    const float* bk = sm.blk;
    const V3 bx = v3(bk[BK_POS], bk[BK_POS + 1], bk[BK_POS + 2]), bv = v3(bk[BK_V], bk[BK_V + 1], bk[BK_V + 2]), bw = v3(bk[BK_W], bk[BK_W + 1], bk[BK_W + 2]);
    const V3 rel = tip - bx, rv = tv - bv, rw = tw - bw;
    float* row = io.obs + (size_t)env * (D::W + 2 * jo);
    float* obs = row + jo; float* pol = row + D::O + 2 * jo; float* ag = pol + D::P; float* dg = ag + D::G;
    obs[0] = tip.x; obs[1] = tip.y; obs[2] = tip.z; obs[3] = bx.x; obs[4] = bx.y; obs[5] = bx.z; obs[6] = closeness;
    obs[7] = rel.x; obs[8] = rel.y; obs[9] = rel.z; obs[10] = tv.x; obs[11] = tv.y; obs[12] = tv.z; obs[13] = finger_vel;
    obs[14] = rv.x; obs[15] = rv.y; obs[16] = rv.z; obs[17] = rw.x; obs[18] = rw.y; obs[19] = rw.z;
    pol[0] = tip.x; pol[1] = tip.y; pol[2] = tip.z; pol[3] = closeness; pol[4] = rel.x; pol[5] = rel.y; pol[6] = rel.z;
    ag[0] = bx.x; ag[1] = bx.y; ag[2] = bx.z;
    const float* goal = s + (size_t)(ST_BLK + 13) * B;
    const float g0 = goal[0], g1 = goal[B], g2 = goal[2 * B];
    dg[0] = g0; dg[1] = g1; dg[2] = g2;
    const float dx = bx.x - g0, dy = bx.y - g1, dz = bx.z - g2;
    const float dist = sqrtf(dx * dx + dy * dy + dz * dz);
#pragma unroll
    for (int k = 0; k < 3; k++) s[(ST_EE + k) * B] = ee[k];
    float* el = s + (size_t)(D::STATE - 1) * B;
    const int elapsed = (int)(*el) + 1;
    *el = (float)elapsed;
    const bool na = dist > io.thr;
    io.reward[env] = io.binary ? -(na ? 1.0f : 0.0f) : -dist;
    io.success[env] = na ? 0 : 1;
    io.done[env] = elapsed >= io.max_steps ? 1 : 0;
  }
  gather_push(g, io, env);
}

// ---- multi-block environments: one env.step() -------------------------------------------------------------------
// BlockStack (4-column Cartesian action, grasping) / BlockRearrange (3 columns) with NBLK blocks, including the grip-informed goal
// and the task-decomposition / curriculum sub-goals (kuka_multi_step_base_env.py:255-345, the same assembly as
// write_obs<3, NBLK> of the thread-per-env kernel).  Verified against the oracle on the CPU (tests/emu); not yet
// instantiated in libpmg.so -- dispatch and GPU measurement are the next step.
template <int NBLK>
__device__ void step_env_multi(const Grp& g, EnvSmemT<NBLK>& sm, const float* lane_consts, const StepIO& io, int env) {
  using SM = EnvSmemT<NBLK>;
  using D = Dims<3, NBLK>;
  const int lane = g.lane;
  const bool arm = lane < 7, hand = lane == 7;
  const size_t B = io.tile;  // distance between consecutive words of this environment (StepIO::tile)
  float* s = io.state + state_off(io, env);
  Lane L;
  L.lc = lane_consts + lane * LC_W;
  L.dof0 = lane;
  L.q0 = s[(ST_Q + lane) * B]; L.qd0 = s[(ST_QD + lane) * B]; L.mt0 = s[(ST_MT + lane) * B]; L.mi0 = s[(ST_MI + lane) * B];
  L.q1 = hand ? s[(ST_Q + 8) * B] : 0.0f; L.qd1 = hand ? s[(ST_QD + 8) * B] : 0.0f;
  L.mt1 = hand ? s[(ST_MT + 8) * B] : 0.0f; L.mi1 = hand ? s[(ST_MI + 8) * B] : 0.0f;
  L.dtau0 = L.dtau1 = 0.0f;
  load_manifolds<SM::NPAIRS>(g, sm.man, io.manifold + man_off(io, env), B);
  for (int w = lane; w < 13 * NBLK; w += GL) sm.blk[24 * (w / 13) + w % 13] = s[(size_t)(ST_BLK + w) * B];
  if (lane == 0) {
    sm.blk[23] = 0.0f;  // contact points dropped because the row pool was full
    sm.spill = io.row_spill + (size_t)env * SM::SPILL_WORDS;
  }
  // ---- Kuka.apply_action (kuka.py:167-222) ----
  const float* act = io.action + (size_t)env * io.adim;  // 4 columns (BlockStack: grasping) or 3 (BlockRearrange)
  if (hand && io.grasp) {  // kuka.py:169-172
    const float grip = (act[3] + 1.0f) * (GRIPPER_ABS_LIMIT / 2);
    L.mt0 = L.mt1 = grip; L.mi0 = L.mi1 = FINGER_FORCE * OUTER_DT;
  }
  const float lo[3] = {-0.67f, -0.20f, 0.175f}, hi[3] = {-0.37f, 0.20f, 0.55f};  // kuka.py:40-41
  float ee[3];
#pragma unroll
  for (int k = 0; k < 3; k++) ee[k] = fminf(fmaxf(s[(ST_EE + k) * B] + act[k] * 0.01f, lo[k]), hi[k]);
  {
    const float tq[4] = {0.f, -1.f, 0.f, 0.f};  // kuka.py:42
    const float qik = inverse_kinematics(g, L, arm ? L.q0 : 0.0f, v3(ee[0], ee[1], ee[2]), tq);
    if (arm) { L.mt0 = qik; L.mi0 = ARM_FORCE * OUTER_DT; }
  }
  g.sync();
  // ---- 5 x stepSimulation (kuka.py:223-225), each 20 substeps of 2 ms ----
  for (int call = 0; call < CALLS_PER_ENV_STEP; call++) {
    L.dtau0 = -L.lc[LC_DAMP] * L.qd0;
    L.dtau1 = 0.0f;
    for (int sub = 0; sub < SUBSTEPS_PER_CALL; sub++) { g.block_sync(); substep(g, sm, L); }
  }
  // ---- observation, reward, flags ----
  M3 R; V3 p;
  chain_fk(g, L, arm ? L.q0 : 0.0f, R, p);
  const V3 a = arm ? col(R, 2) : v3(0, 0, 0);
  const V3 aq = L.qd0 * a;
  const V3 w = g.scan(aq), wp = w - aq;
  V3 pprev = g.up(p, 1);
  if (lane == 0) pprev = v3(0, 0, 0);
  const V3 vo = g.scan(cross(wp, p - pprev));
  s[(ST_Q + lane) * B] = L.q0; s[(ST_QD + lane) * B] = L.qd0; s[(ST_MT + lane) * B] = L.mt0; s[(ST_MI + lane) * B] = L.mi0;
  if (hand) { s[(ST_Q + 8) * B] = L.q1; s[(ST_QD + 8) * B] = L.qd1; s[(ST_MT + 8) * B] = L.mt1; s[(ST_MI + 8) * B] = L.mi1; }
  g.sync();  // the block states of the last substep
  store_manifolds<SM::NPAIRS>(g, sm.man, io.manifold + man_off(io, env), B);
  for (int wd = lane; wd < 13 * NBLK; wd += GL) s[(size_t)(ST_BLK + wd) * B] = sm.blk[24 * (wd / 13) + wd % 13];
  if (lane == 0 && sm.blk[23] > 0.0f && io.overflow) atomicAdd(io.overflow, (int)sm.blk[23]);
  const float t[3] = PMG_TIP_OFFSET;
  const V3 tip_own = p + mul(R, v3(t[0], t[1], t[2]));
  const V3 tip = g.shfl(tip_own, PMG_BODY_LINK7);
  const V3 tv = g.shfl(vo + cross(w, tip_own - p), PMG_BODY_LINK7), tw = g.shfl(w, PMG_BODY_LINK7);
  if (hand) {
    const float t1[3] = PMG_TAB1_OFFSET, t2[3] = PMG_TAB2_OFFSET;
    const V3 ay = col(R, 1);  // R = gripper-base frame on this lane; fingers slide along -/+ its y axis
    const V3 j1 = v3(c_jxyz[PMG_BODY_FINGER1][0], c_jxyz[PMG_BODY_FINGER1][1], c_jxyz[PMG_BODY_FINGER1][2]);
    const V3 j2 = v3(c_jxyz[PMG_BODY_FINGER2][0], c_jxyz[PMG_BODY_FINGER2][1], c_jxyz[PMG_BODY_FINGER2][2]);
    const V3 tab1 = mul(R, j1 + v3(t1[0], t1[1], t1[2])) - L.q0 * ay;
    const V3 tab2 = mul(R, j2 + v3(t2[0], t2[1], t2[2])) + L.q1 * ay;
    const float closeness = io.grasp ? norm(tab1 - tab2) : 0.0f;   // kuka.py:245-246: [0.0] without grasping
    const float finger_vel = io.grasp ? -(cross(w, tab1) - L.qd0 * ay).y : 0.0f;
    const int G = io.goal_dim;  // 3 NBLK, + 4 with the grip-informed goal
    float* row = io.obs + (size_t)env * io.row_width;
    float* obs = row; float* pol = row + D::O; float* ag = pol + D::P; float* dg = ag + G;
    obs[0] = tip.x; obs[1] = tip.y; obs[2] = tip.z; obs[3] = closeness; obs[4] = tv.x; obs[5] = tv.y; obs[6] = tv.z; obs[7] = finger_vel;
    pol[0] = tip.x; pol[1] = tip.y; pol[2] = tip.z; pol[3] = closeness;
#pragma unroll 1
    for (int n = 0; n < NBLK; n++) {
      const float* bk = sm.blk + 24 * n;
      float* bs = obs + 8 + 16 * n;
      const V3 bx = v3(bk[BK_POS], bk[BK_POS + 1], bk[BK_POS + 2]), bv = v3(bk[BK_V], bk[BK_V + 1], bk[BK_V + 2]), bw = v3(bk[BK_W], bk[BK_W + 1], bk[BK_W + 2]);
      const V3 rel = tip - bx, rv = tv - bv, rw = tw - bw;
      bs[0] = bx.x; bs[1] = bx.y; bs[2] = bx.z; bs[3] = rel.x; bs[4] = rel.y; bs[5] = rel.z;
      bs[6] = bk[BK_QUAT]; bs[7] = bk[BK_QUAT + 1]; bs[8] = bk[BK_QUAT + 2]; bs[9] = bk[BK_QUAT + 3];
      bs[10] = rv.x; bs[11] = rv.y; bs[12] = rv.z; bs[13] = rw.x; bs[14] = rw.y; bs[15] = rw.z;
      pol[4 + 3 * n] = rel.x; pol[5 + 3 * n] = rel.y; pol[6 + 3 * n] = rel.z;
      ag[3 * n] = bx.x; ag[3 * n + 1] = bx.y; ag[3 * n + 2] = bx.z;
    }
    if (io.grip_goal) { ag[3 * NBLK] = tip.x; ag[3 * NBLK + 1] = tip.y; ag[3 * NBLK + 2] = tip.z; ag[3 * NBLK + 3] = closeness; }
    const float* goal = s + (size_t)(ST_BLK + 13 * NBLK) * B;
    for (int k = 0; k < G; k++) dg[k] = goal[(size_t)k * B];
    if (io.td == 2) {  // BlockRearrange curriculum: only the blocks of the episode's mask have targets (write_obs in pmg_capi.cu)
      const int moved = (int)goal[(size_t)G * B];
#pragma unroll 1
      for (int n = 0; n < NBLK; n++)
        if (!((moved >> n) & 1)) { const float* bk = sm.blk + 24 * n; dg[3 * n] = bk[BK_POS]; dg[3 * n + 1] = bk[BK_POS + 1]; dg[3 * n + 2] = bk[BK_POS + 2]; }
    } else if (io.td) {  // sub-goal rebuilt from the current block positions (see write_obs in pmg_capi.cu)
      const int nsub = io.grip_goal ? 2 * NBLK : NBLK;
      int ind = (int)goal[(size_t)G * B];
      if (ind < 0) ind += nsub;
      const int k = io.grip_goal ? ind >> 1 : ind;
      const bool place = io.grip_goal ? (ind & 1) != 0 : true;
#pragma unroll 1
      for (int n = 0; n < NBLK; n++) {
        const float* bk = sm.blk + 24 * n;
        const int level = (int)floorf((dg[3 * n + 2] - BLOCK_SPAWN_Z) * (1.0f / 0.03f) + 0.5f);
        const bool at_target = place ? level <= k : level < k;
        if (io.grip_goal && level == k) {
          dg[3 * NBLK] = place ? dg[3 * n] : bk[BK_POS]; dg[3 * NBLK + 1] = place ? dg[3 * n + 1] : bk[BK_POS + 1];
          dg[3 * NBLK + 2] = place ? dg[3 * n + 2] : bk[BK_POS + 2];
        }
        if (!at_target) { dg[3 * n] = bk[BK_POS]; dg[3 * n + 1] = bk[BK_POS + 1]; dg[3 * n + 2] = bk[BK_POS + 2]; }
      }
    }
    for (int k = 0; k < D::O + D::P; k++) row[k] = fminf(fmaxf(row[k], -5.0f), 5.0f);  // np.clip of observation and policy_state
    float d2 = 0.0f;
    for (int k = 0; k < G; k++) { const float d = ag[k] - dg[k]; d2 += d * d; }
    const float dist = sqrtf(d2);
#pragma unroll
    for (int k = 0; k < 3; k++) s[(ST_EE + k) * B] = ee[k];
    float* el = s + (size_t)(io.state_words - 1) * B;
    const int elapsed = (int)(*el) + 1;
    *el = (float)elapsed;
    const bool na = dist > io.thr;
    io.reward[env] = io.binary ? -(na ? 1.0f : 0.0f) : -dist;
    io.success[env] = na ? 0 : 1;
    io.done[env] = elapsed >= io.max_steps ? 1 : 0;
  }
  gather_push(g, io, env);
}

}  // namespace coop
}  // namespace pmg
