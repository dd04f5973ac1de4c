// pmg_sim.cuh -- one environment's substep / env-step on top of pmg_physics.cuh.
//
// Persistent per-env state lives in two struct-of-arrays buffers in HBM:
//   state[word][env]   : q[9] qd[9] ee_target[3] rest_pose[7] motor_target[9] motor_max_impulse[9]
//                        | per block pos[3] quat[4] linvel[3] angvel[3] | desired_goal[G] | elapsed
//   manifold[word][env]: per collision pair: count, then 4 x (localA[3] localB[3] normalB[3] dist)
// Thread t of a warp touches word w at address (w*batch + env), so every access of a warp is one
// or two 128-byte lines.
#pragma once

#include <stddef.h>

#include "pmg_physics.cuh"

namespace pmg {

constexpr int MAN_WORDS = 41;  // per pair
constexpr int ST_Q = 0, ST_QD = 9, ST_EE = 18, ST_REST = 21, ST_MT = 28, ST_MI = 37, ST_BLK = 46;

// finger-table (2), per block table / floor / finger1 / finger2 (4 each), block-block, and -- scenes with two or more
// blocks, where a stack reaches it -- the gripper-base cylinder against every block
__host__ __device__ constexpr int num_pairs(int nblk) { return 2 + 4 * nblk + nblk * (nblk - 1) / 2 + (nblk >= 2 ? nblk : 0); }
__host__ __device__ constexpr int max_points(int nblk) { return nblk <= 1 ? 4 * num_pairs(nblk) : 48; }
__host__ __device__ constexpr int max_robot_points(int nblk) { return nblk == 0 ? 8 : 16; }

// geometry endpoints of a collision pair
enum GeomKind { G_TABLE = 0, G_FLOOR = 1, G_FINGER1 = 2, G_FINGER2 = 3, G_BLOCK = 4, G_GBASE = 5 /* gripper-base cylinder: (r, r, h) */ };
struct PairInfo { int ka, kb, ia, ib; };  // kinds and block indices (for G_BLOCK)

template <int NBLK>
__device__ __forceinline__ PairInfo pair_info(int k) {
  PairInfo p; p.ia = p.ib = 0;
  if (k < 2) { p.ka = G_FINGER1 + k; p.kb = G_TABLE; return p; }
  k -= 2;
  if (k < 4 * NBLK) {
    int i = k >> 2, w = k & 3;
    p.kb = G_BLOCK; p.ib = i;
    p.ka = w == 0 ? G_TABLE : (w == 1 ? G_FLOOR : (w == 2 ? G_FINGER1 : G_FINGER2));
    return p;
  }
  k -= 4 * NBLK;
  if (k >= NBLK * (NBLK - 1) / 2) {  // gripper base (a robot end, like the fingers) against block k
    p.ka = G_GBASE; p.kb = G_BLOCK; p.ib = k - NBLK * (NBLK - 1) / 2;
    return p;
  }
  int i = 0;
  while (k >= NBLK - 1 - i) { k -= NBLK - 1 - i; i++; }
  p.ka = G_BLOCK; p.kb = G_BLOCK; p.ia = i; p.ib = i + 1 + k;
  return p;
}
__device__ __forceinline__ bool geom_robot(int kind) { return kind == G_FINGER1 || kind == G_FINGER2 || kind == G_GBASE; }

// puck: the Slide scene (long low-friction table, the block is a cylinder about its z axis: half = (r, r, h))
__device__ __forceinline__ V3 geom_half(int kind, bool puck = false) {
  const float th[3] = PMG_TABLE_HALF, lh[3] = PMG_LONG_TABLE_HALF, fh[3] = PMG_FLOOR_HALF, gh[3] = PMG_FINGER_HALF;
  if (kind == G_TABLE) return puck ? v3(lh[0], lh[1], lh[2]) : v3(th[0], th[1], th[2]);
  if (kind == G_FLOOR) return v3(fh[0], fh[1], fh[2]);
  if (kind == G_BLOCK) return puck ? v3((float)PMG_PUCK_RADIUS, (float)PMG_PUCK_RADIUS, (float)PMG_PUCK_HALF_LEN) : v3(BLOCK_HALF, BLOCK_HALF, BLOCK_HALF);
  if (kind == G_GBASE) return v3((float)PMG_GBASE_RADIUS, (float)PMG_GBASE_RADIUS, (float)PMG_GBASE_HALF_LEN);
  return v3(gh[0], gh[1], gh[2]);
}
__device__ __forceinline__ V3 table_center(bool puck = false) {
  const float tc[3] = PMG_TABLE_CENTER, lc[3] = PMG_LONG_TABLE_CENTER;
  return puck ? v3(lc[0], lc[1], lc[2]) : v3(tc[0], tc[1], tc[2]);
}
__device__ __forceinline__ float geom_friction(int kind, bool puck = false) {
  if (kind == G_TABLE) return puck ? (float)PMG_LONG_TABLE_FRICTION : (float)PMG_TABLE_FRICTION;
  if (kind == G_BLOCK && puck) return (float)PMG_PUCK_FRICTION;
  if (kind == G_FLOOR) return (float)PMG_FLOOR_FRICTION;
  if (kind == G_BLOCK) return (float)PMG_BLOCK_FRICTION;
  if (kind == G_GBASE) return (float)PMG_GBASE_FRICTION;
  return (float)PMG_FINGER_FRICTION;
}

// Where one environment's persistent manifolds live: global memory ([word][env], stride = batch) for the
// thread-per-env kernels, a shared-memory copy (stride 1) for the lane-cooperative kernel (pmg_coop.cuh).
struct ManRef {
  float* man;      // this env's first manifold word
  size_t stride;   // distance between consecutive words
  __device__ __forceinline__ float& mw(int pair, int w) { return man[(size_t)(pair * MAN_WORDS + w) * stride]; }
};

// Manifold-local coordinates of a static box refer to the centre of its TOP FACE, not to its centre: contact
// points on the table are then small numbers (see box_box: the fp32 rounding of a 0.08 m coordinate would
// otherwise show up as spin of resting blocks).  Dynamic bodies are small; theirs refer to the body centre.
__device__ __forceinline__ V3 geom_anchor(int kind) {
  const float th[3] = PMG_TABLE_HALF, fh[3] = PMG_FLOOR_HALF;
  if (kind == G_TABLE) return v3(0.0f, 0.0f, th[2]);
  if (kind == G_FLOOR) return v3(0.0f, 0.0f, fh[2]);
  return v3(0.0f, 0.0f, 0.0f);
}
__device__ __forceinline__ bool geom_static(int kind) { return kind == G_TABLE || kind == G_FLOOR; }

template <int NBLK>
struct Env : ManRef {
  static constexpr int NBA = NBLK > 0 ? NBLK : 1;
  float q[ND], qd[ND], mt[ND], mi[ND], dtau[ND];
  V3 bpos[NBA], bv[NBA], bw[NBA];
  float bquat[NBA][4];
  M3 bR[NBA];
  int overflow;
};

template <int NBLK>
__device__ __forceinline__ void geom_pose(const Env<NBLK>& e, const Frames& f, int kind, int idx, V3& p, M3& R) {
  const float tc[3] = PMG_TABLE_CENTER, fc[3] = PMG_FLOOR_CENTER;
  if (kind == G_TABLE) { p = v3(tc[0], tc[1], tc[2]); R = m3_identity(); }
  else if (kind == G_FLOOR) { p = v3(fc[0], fc[1], fc[2]); R = m3_identity(); }
  else if (kind == G_FINGER1) { p = f.p[PMG_BODY_FINGER1]; R = f.R[PMG_BODY_FINGER1]; }
  else if (kind == G_FINGER2) { p = f.p[PMG_BODY_FINGER2]; R = f.R[PMG_BODY_FINGER2]; }
  else if (kind == G_GBASE) { p = f.p[PMG_BODY_GBASE]; R = f.R[PMG_BODY_GBASE]; }
  else { p = e.bpos[idx]; R = e.bR[idx]; }
}

// ---- collision detection + persistent manifolds (btPersistentManifold semantics) -----------
__device__ void manifold_add(ManRef& e, int k, float thr, V3 lA, V3 lB, V3 nB, float dist) {
  int n = __float_as_int(e.mw(k, 0));
  float shortest = thr * thr;
  int nearest = -1;
  for (int i = 0; i < n; i++) {
    V3 c = v3(e.mw(k, 1 + 10 * i), e.mw(k, 2 + 10 * i), e.mw(k, 3 + 10 * i));
    V3 d = c - lA;
    float dd = dot(d, d);
    if (dd < shortest) { shortest = dd; nearest = i; }
  }
  int slot = nearest;
  if (slot < 0) {
    if (n < 4) { slot = n; e.mw(k, 0) = __int_as_float(n + 1); }
    else {
      V3 c[4]; float dd[4];
      for (int i = 0; i < 4; i++) { c[i] = v3(e.mw(k, 1 + 10 * i), e.mw(k, 2 + 10 * i), e.mw(k, 3 + 10 * i)); dd[i] = e.mw(k, 10 + 10 * i); }
      int deepest = -1; float maxpen = dist;
      for (int i = 0; i < 4; i++) if (dd[i] < maxpen) { deepest = i; maxpen = dd[i]; }
      float res[4] = {0, 0, 0, 0};
      V3 x;
      if (deepest != 0) { x = cross(lA - c[1], c[3] - c[2]); res[0] = dot(x, x); }
      if (deepest != 1) { x = cross(lA - c[0], c[3] - c[2]); res[1] = dot(x, x); }
      if (deepest != 2) { x = cross(lA - c[0], c[3] - c[1]); res[2] = dot(x, x); }
      if (deepest != 3) { x = cross(lA - c[0], c[2] - c[1]); res[3] = dot(x, x); }
      slot = 0; float best = fabsf(res[0]);
      for (int i = 1; i < 4; i++) if (fabsf(res[i]) > best) { best = fabsf(res[i]); slot = i; }
    }
  }
  int o = 1 + 10 * slot;
  e.mw(k, o) = lA.x; e.mw(k, o + 1) = lA.y; e.mw(k, o + 2) = lA.z;
  e.mw(k, o + 3) = lB.x; e.mw(k, o + 4) = lB.y; e.mw(k, o + 5) = lB.z;
  e.mw(k, o + 6) = nB.x; e.mw(k, o + 7) = nB.y; e.mw(k, o + 8) = nB.z;
  e.mw(k, o + 9) = dist;
}

// btPersistentManifold::refreshContactPoints: drop points that separated or drifted, update distances
__device__ void manifold_refresh(ManRef& e, int k, float thr, V3 pa, const M3& Ra, V3 pb, const M3& Rb) {
  int n = __float_as_int(e.mw(k, 0));
  for (int i = n - 1; i >= 0; i--) {
    int o = 1 + 10 * i;
    V3 lA = v3(e.mw(k, o), e.mw(k, o + 1), e.mw(k, o + 2)), lB = v3(e.mw(k, o + 3), e.mw(k, o + 4), e.mw(k, o + 5));
    V3 nB = v3(e.mw(k, o + 6), e.mw(k, o + 7), e.mw(k, o + 8));
    V3 wa = mul(Ra, lA) + pa, wb = mul(Rb, lB) + pb;
    float dist = dot(wa - wb, nB);
    bool drop = dist > thr;
    if (!drop) {
      V3 pd = wb - (wa - dist * nB);
      drop = dot(pd, pd) > thr * thr;
    }
    if (drop) {
      int last = n - 1;
      if (i != last) for (int w = 0; w < 10; w++) e.mw(k, o + w) = e.mw(k, 1 + 10 * last + w);
      n--;
    } else e.mw(k, o + 9) = dist;
  }
  if (n != __float_as_int(e.mw(k, 0))) e.mw(k, 0) = __int_as_float(n);
}

#ifdef PMG_COOP_TIMING
__device__ unsigned long long g_coop_cycles[16];
#endif

// One collision pair: broadphase (world AABBs grown by the margin), box-box narrowphase into the
// persistent manifold, refresh.  Boxes given by centre, orientation, half extents; aA / aB are the anchors
// (geom_anchor, world axes: static boxes are axis aligned) the manifold-local coordinates refer to.  The
// narrowphase and the manifold run in coordinates relative to the static box's anchor (else to B's centre),
// so that penetration depths are differences of small numbers.
// cylB: body B is the Slide puck (a cylinder about its z axis, hb = (r, r, h)): box_cyl instead of box_box.
// cylA: body A is the gripper-base cylinder (ha = (r, r, h)): box_cyl with the roles exchanged (cyl_box).
// b_static: body B is the static box (the finger-table pairs); with a_static it selects box_box's static fast path.
__device__ void collide_pair(ManRef& e, int k, V3 pa, const M3& Ra, V3 ha, V3 aA, bool a_static, V3 pb, const M3& Rb, V3 hb, V3 aB, BoxScratch& scr,
                             bool cylB = false, bool b_static = false, bool cylA = false) {
  V3 d = pa - pb;
  float ex = fabsf(Ra.r0.x) * ha.x + fabsf(Ra.r0.y) * ha.y + fabsf(Ra.r0.z) * ha.z + fabsf(Rb.r0.x) * hb.x + fabsf(Rb.r0.y) * hb.y + fabsf(Rb.r0.z) * hb.z + 2 * BROADPHASE_MARGIN;
  float ey = fabsf(Ra.r1.x) * ha.x + fabsf(Ra.r1.y) * ha.y + fabsf(Ra.r1.z) * ha.z + fabsf(Rb.r1.x) * hb.x + fabsf(Rb.r1.y) * hb.y + fabsf(Rb.r1.z) * hb.z + 2 * BROADPHASE_MARGIN;
  float ez = fabsf(Ra.r2.x) * ha.x + fabsf(Ra.r2.y) * ha.y + fabsf(Ra.r2.z) * ha.z + fabsf(Rb.r2.x) * hb.x + fabsf(Rb.r2.y) * hb.y + fabsf(Rb.r2.z) * hb.z + 2 * BROADPHASE_MARGIN;
  if (fabsf(d.x) > ex || fabsf(d.y) > ey || fabsf(d.z) > ez) {
    if (__float_as_int(e.mw(k, 0)) != 0) e.mw(k, 0) = __int_as_float(0);
    return;
  }
  float thr = BREAKING_THRESHOLD_FACTOR * fminf(norm(ha), norm(hb));
  const V3 O = a_static ? pa + aA : pb + aB;            // origin of the relative coordinates
  const V3 a0 = pa - O, b0 = pb - O;                     // box centres
  const V3 a1 = a0 + aA, b1 = b0 + aB;                   // anchors of the manifold-local coordinates
  const Contact* c = scr.out;
#ifdef PMG_COOP_TIMING
  const long long t0 = clock64();
#endif
  int nc;
  if (cylA) nc = cyl_box(a0, Ra, ha.x, ha.z, b0, Rb, hb, scr.out);  // the gripper base against a block
  else nc = cylB ? box_cyl(a0, Ra, ha, b0, Rb, hb.x, hb.z, scr.out) : box_box(a0, Ra, ha, b0, Rb, hb, scr, a_static ? 1 : (b_static ? 2 : 0));
#ifdef PMG_COOP_TIMING
  const long long t1 = clock64();
#endif
  for (int i = 0; i < nc; i++) {
    V3 wa = c[i].pB + c[i].dist * c[i].nB;
    manifold_add(e, k, thr, mulT(Ra, wa - a1), mulT(Rb, c[i].pB - b1), c[i].nB, c[i].dist);
  }
#ifdef PMG_COOP_TIMING
  const long long t2 = clock64();
#endif
  manifold_refresh(e, k, thr, a1, Ra, b1, Rb);
#ifdef PMG_COOP_TIMING
  if (k == 0 && nc > 0) {
    atomicAdd(&g_coop_cycles[8], (unsigned long long)(t1 - t0)); atomicAdd(&g_coop_cycles[9], (unsigned long long)(t2 - t1));
    atomicAdd(&g_coop_cycles[10], (unsigned long long)(clock64() - t2)); atomicAdd(&g_coop_cycles[11], 1ull);
  }
#endif
}

template <int NBLK>
__device__ void collide(Env<NBLK>& e, const Frames& f) {
#pragma unroll
  for (int b = 0; b < NBLK; b++) e.bR[b] = quat_to_m3(e.bquat[b][0], e.bquat[b][1], e.bquat[b][2], e.bquat[b][3]);
  constexpr int NP = num_pairs(NBLK);
  for (int k = 0; k < NP; k++) {
    PairInfo pi = pair_info<NBLK>(k);
    V3 pa, pb; M3 Ra, Rb;
    geom_pose(e, f, pi.ka, pi.ia, pa, Ra);
    geom_pose(e, f, pi.kb, pi.ib, pb, Rb);
    V3 ha = geom_half(pi.ka), hb = geom_half(pi.kb);
    // broadphase: world AABBs grown by gContactBreakingThreshold
    V3 d = pa - pb;
    float ex = fabsf(Ra.r0.x) * ha.x + fabsf(Ra.r0.y) * ha.y + fabsf(Ra.r0.z) * ha.z + fabsf(Rb.r0.x) * hb.x + fabsf(Rb.r0.y) * hb.y + fabsf(Rb.r0.z) * hb.z + 2 * BROADPHASE_MARGIN;
    float ey = fabsf(Ra.r1.x) * ha.x + fabsf(Ra.r1.y) * ha.y + fabsf(Ra.r1.z) * ha.z + fabsf(Rb.r1.x) * hb.x + fabsf(Rb.r1.y) * hb.y + fabsf(Rb.r1.z) * hb.z + 2 * BROADPHASE_MARGIN;
    float ez = fabsf(Ra.r2.x) * ha.x + fabsf(Ra.r2.y) * ha.y + fabsf(Ra.r2.z) * ha.z + fabsf(Rb.r2.x) * hb.x + fabsf(Rb.r2.y) * hb.y + fabsf(Rb.r2.z) * hb.z + 2 * BROADPHASE_MARGIN;
    if (fabsf(d.x) > ex || fabsf(d.y) > ey || fabsf(d.z) > ez) {
      if (__float_as_int(e.mw(k, 0)) != 0) e.mw(k, 0) = __int_as_float(0);
      continue;
    }
    float thr = BREAKING_THRESHOLD_FACTOR * fminf(norm(ha), norm(hb));
    // narrowphase and manifold in coordinates relative to the static box's anchor (else B's centre), see collide_pair
    const V3 aA = geom_anchor(pi.ka), aB = geom_anchor(pi.kb);
    const V3 O = geom_static(pi.ka) ? pa + aA : pb + aB;
    const V3 a0 = pa - O, b0 = pb - O, a1 = a0 + aA, b1 = b0 + aB;
    BoxScratch scr;
    const Contact* c = scr.out;
    int nc = pi.ka == G_GBASE ? cyl_box(a0, Ra, ha.x, ha.z, b0, Rb, hb, scr.out)
                              : box_box(a0, Ra, ha, b0, Rb, hb, scr, geom_static(pi.ka) ? 1 : (geom_static(pi.kb) ? 2 : 0));
    for (int i = 0; i < nc; i++) {
      V3 wa = c[i].pB + c[i].dist * c[i].nB;
      manifold_add(e, k, thr, mulT(Ra, wa - a1), mulT(Rb, c[i].pB - b1), c[i].nB, c[i].dist);
    }
    manifold_refresh(e, k, thr, a1, Ra, b1, Rb);
  }
}

// ---- constraint rows + projected Gauss-Seidel ------------------------------------------------
// Non-contact rows.  Bullet visits the 18 constraints in its sorted order c_nc_order: with the
// generated order that is the 9 joint motors (dofs 2,3,0,1,4,7,8,5,6) followed by the 9 joint-limit
// constraints in the same dof order, each limit constraint contributing its lower-bound row (+1)
// and then its upper-bound row (-1) when violated.  The motor rows are always active and are kept
// branch-free in registers; limit rows are rare (the closed jaws sit on their upper limit) and go
// through a small rolled loop.  The whole solver stays a few hundred instructions: the step kernel
// is bound by instruction fetch and dependent-issue latency, not by FLOPs (profiles/).
__host__ __device__ constexpr int nc_order(int i) {
  constexpr int o[2 * ND] = PMG_NONCONTACT_ORDER;
  return o[i];
}
__host__ __device__ constexpr bool nc_order_is_motors_then_limits() {
  for (int i = 0; i < ND; i++)
    if (nc_order(i) < ND || nc_order(ND + i) >= ND || nc_order(i) - ND != nc_order(ND + i)) return false;
  return true;
}
static_assert(nc_order_is_motors_then_limits(), "solver layout assumes motors first, then limits, same dof order");
__host__ __device__ constexpr int nc_dof(int k) { return nc_order(k) - ND; }  // dof of the k-th motor / limit constraint

struct NcRows {
  float rhs[ND], lim[ND], app[ND], dinv[ND];  // motor rows, indexed by visit position k; dinv = 1 / Minv[d][d]
  float lrhs[2 * ND], lapp[2 * ND];  // limit rows, slot 2*k + side
  unsigned lactive;                  // bit per limit slot
};

__device__ __forceinline__ void nc_setup(NcRows& nc, const float* q, const float* qd, const float* mt, const float* mi, const float (*Minv)[ND]) {
  nc.lactive = 0u;
#pragma unroll
  for (int k = 0; k < ND; k++) {
    const int d = nc_dof(k);
    const float dinv = 1.0f / Minv[d][d];
    // btMultiBodyJointMotor in POSITION_CONTROL (kuka.py:282-301): velocity target from the position error
    float target_vel = MOTOR_KP * (mt[d] - q[d]) * INV_DT + qd[d] + MOTOR_KD * (0.0f - qd[d]);
    nc.rhs[k] = (target_vel - qd[d]) * dinv; nc.lim[k] = mi[d]; nc.app[k] = 0.0f; nc.dinv[k] = dinv;
    // btMultiBodyJointLimitConstraint: row 0 lower bound (+), row 1 upper bound (-), only when violated
    float pen0 = q[d] - c_dof_lower[d], pen1 = c_dof_upper[d] - q[d];
    if (!(pen0 > 0.0f)) {
      float pos_err = pen0 > SPLIT_IMPULSE_PEN_THRESHOLD ? -pen0 * CONTACT_ERP * INV_DT : 0.0f;
      nc.lactive |= 1u << (2 * k);
      nc.lrhs[2 * k] = (pos_err - qd[d]) * dinv; nc.lapp[2 * k] = 0.0f;
    }
    if (!(pen1 > 0.0f)) {
      float pos_err = pen1 > SPLIT_IMPULSE_PEN_THRESHOLD ? -pen1 * CONTACT_ERP * INV_DT : 0.0f;
      nc.lactive |= 1u << (2 * k + 1);
      nc.lrhs[2 * k + 1] = (pos_err + qd[d]) * dinv; nc.lapp[2 * k + 1] = 0.0f;
    }
  }
}

template <int K>
__device__ __forceinline__ void nc_motor_row(NcRows& nc, const float (*Minv)[ND], float* dqd, float& res) {
  constexpr int d = nc_dof(K);
  const float mdd = Minv[d][d];
  float dl = nc.rhs[K] - dqd[d] * nc.dinv[K];
  float sum = fminf(fmaxf(nc.app[K] + dl, -nc.lim[K]), nc.lim[K]);
  dl = sum - nc.app[K];
  nc.app[K] = sum;
#pragma unroll
  for (int r = 0; r < ND; r++) dqd[r] += Minv[d][r] * dl;
  float rr = dl * mdd;
  res = fmaxf(res, rr * rr);
}

__device__ __forceinline__ void nc_limit_rows(NcRows& nc, bool forward, const float (*Minv)[ND], float* dqd, float& res) {
  unsigned todo = nc.lactive;
#pragma unroll 1
  while (todo) {
    const int s = forward ? __ffs(todo) - 1 : 31 - __clz(todo);
    todo &= ~(1u << s);
    const int d = c_nc_order[ND + (s >> 1)];
    const float sign = (s & 1) ? -1.0f : 1.0f;
    const float* col = Minv[d];  // symmetric: column d == row d
    float cur = 0.0f;
#pragma unroll
    for (int r = 0; r < ND; r++) cur = r == d ? dqd[r] : cur;
    float dl = nc.lrhs[s] - sign * cur * (1.0f / col[d]);
    float sum = fminf(fmaxf(nc.lapp[s] + dl, 0.0f), LIMIT_MAX_IMPULSE);
    dl = sum - nc.lapp[s];
    nc.lapp[s] = sum;
    const float sdl = sign * dl;
#pragma unroll
    for (int r = 0; r < ND; r++) dqd[r] += col[r] * sdl;
    float rr = dl * col[d];
    res = fmaxf(res, rr * rr);
  }
}

// One Gauss-Seidel sweep over the non-contact rows: backwards on even iterations, forwards on odd.
__device__ __forceinline__ void nc_sweep(NcRows& nc, bool forward, const float (*Minv)[ND], float* dqd, float& res) {
  if (forward) {
    nc_motor_row<0>(nc, Minv, dqd, res); nc_motor_row<1>(nc, Minv, dqd, res); nc_motor_row<2>(nc, Minv, dqd, res);
    nc_motor_row<3>(nc, Minv, dqd, res); nc_motor_row<4>(nc, Minv, dqd, res); nc_motor_row<5>(nc, Minv, dqd, res);
    nc_motor_row<6>(nc, Minv, dqd, res); nc_motor_row<7>(nc, Minv, dqd, res); nc_motor_row<8>(nc, Minv, dqd, res);
    nc_limit_rows(nc, true, Minv, dqd, res);
  } else {
    nc_limit_rows(nc, false, Minv, dqd, res);
    nc_motor_row<8>(nc, Minv, dqd, res); nc_motor_row<7>(nc, Minv, dqd, res); nc_motor_row<6>(nc, Minv, dqd, res);
    nc_motor_row<5>(nc, Minv, dqd, res); nc_motor_row<4>(nc, Minv, dqd, res); nc_motor_row<3>(nc, Minv, dqd, res);
    nc_motor_row<2>(nc, Minv, dqd, res); nc_motor_row<1>(nc, Minv, dqd, res); nc_motor_row<0>(nc, Minv, dqd, res);
  }
}

// Contact rows.  One record per row (a contact point has three: normal, tangent 1, tangent 2), sized
// and aligned for 128-bit local-memory accesses:
//   RowRec   : direction d, r_A x d, r_B x d (lever arms about the block centres), rhs, 1/denominator,
//              accumulated impulse.  The cubes have isotropic inertia, so M^-1 J^T of a block endpoint is
//              J scaled by 1/m (linear) and 1/I (angular) and needs no storage.
//   RobotRow : J and M^-1 J^T over the 9 robot dofs, padded to 12, only for points on a finger.
struct __align__(16) RowRec {
  float dx, dy, dz, rhs;
  float ax, ay, az, dinv;   // r_A x d
  float bx, by, bz, app;    // r_B x d
  float denom, mu;          // 1 / dinv; friction coefficient of the pair
  int ends;                 // robot-pool index (bits 0..7, 0xff = none) | block A (8..15) | block B (16..23)
  int pad;
};
__device__ __forceinline__ int rec_rob(int ends) { int v = ends & 0xff; return v == 0xff ? -1 : v; }
__device__ __forceinline__ int rec_blk_a(int ends) { int v = (ends >> 8) & 0xff; return v == 0xff ? -1 : v; }
__device__ __forceinline__ int rec_blk_b(int ends) { int v = (ends >> 16) & 0xff; return v == 0xff ? -1 : v; }
struct __align__(16) RobotRow { float J[12]; float MJ[12]; };

template <int NBLK>
struct ContactRows {
  static constexpr int MP = max_points(NBLK), MR = max_robot_points(NBLK);
  int n, nrob;
  RowRec row[MP][3];
  RobotRow rrow[MR][3];
};

// Block-indexed access that keeps the per-block delta velocities in registers: a select chain over
// the (few) blocks instead of a dynamically indexed local-memory array inside the sequential sweep.
template <int N>
__device__ __forceinline__ V3 blk_get(const V3* v, int b) {
  V3 r = v[0];
#pragma unroll
  for (int k = 1; k < N; k++) { r.x = b == k ? v[k].x : r.x; r.y = b == k ? v[k].y : r.y; r.z = b == k ? v[k].z : r.z; }
  return r;
}
template <int N>
__device__ __forceinline__ void blk_add(V3* v, int b, V3 d) {
#pragma unroll
  for (int k = 0; k < N; k++) { const float m = b == k ? 1.0f : 0.0f; v[k].x += m * d.x; v[k].y += m * d.y; v[k].z += m * d.z; }
}

// velocity of the row's constraint direction for the velocity (or delta-velocity) vq / vlin / vang
template <int NBLK>
__device__ __forceinline__ float row_velocity(const ContactRows<NBLK>& cr, int c, int k, const RowRec& r, const float* vq, const V3* vlin, const V3* vang) {
  float v = 0.0f;
  const int ri = rec_rob(r.ends);
  if (ri >= 0) {
    const RobotRow& rr = cr.rrow[ri][k];
#pragma unroll
    for (int j = 0; j < ND; j++) v += rr.J[j] * vq[j];
  }
  if (NBLK > 0) {
    const int a = rec_blk_a(r.ends), b = rec_blk_b(r.ends);
    if (a >= 0) { V3 l = blk_get<NBLK>(vlin, a), w = blk_get<NBLK>(vang, a); v += r.dx * l.x + r.dy * l.y + r.dz * l.z + r.ax * w.x + r.ay * w.y + r.az * w.z; }
    if (b >= 0) { V3 l = blk_get<NBLK>(vlin, b), w = blk_get<NBLK>(vang, b); v -= r.dx * l.x + r.dy * l.y + r.dz * l.z + r.bx * w.x + r.by * w.y + r.bz * w.z; }
  }
  return v;
}

template <int NBLK>
__device__ __forceinline__ void row_apply(const ContactRows<NBLK>& cr, int c, int k, const RowRec& r, float dl, float* dqd, V3* dlin, V3* dang) {
  const int ri = rec_rob(r.ends);
  if (ri >= 0) {
    const RobotRow& rr = cr.rrow[ri][k];
#pragma unroll
    for (int j = 0; j < ND; j++) dqd[j] += rr.MJ[j] * dl;
  }
  if (NBLK > 0) {
    const int a = rec_blk_a(r.ends), b = rec_blk_b(r.ends);
    const float lm = dl * BLOCK_INV_MASS, li = dl * BLOCK_INV_INERTIA;
    if (a >= 0) { blk_add<NBLK>(dlin, a, v3(lm * r.dx, lm * r.dy, lm * r.dz)); blk_add<NBLK>(dang, a, v3(li * r.ax, li * r.ay, li * r.az)); }
    if (b >= 0) { blk_add<NBLK>(dlin, b, v3(-lm * r.dx, -lm * r.dy, -lm * r.dz)); blk_add<NBLK>(dang, b, v3(-li * r.bx, -li * r.by, -li * r.bz)); }
  }
}

template <int NBLK>
__device__ void solve_constraints(Env<NBLK>& e, const Frames& f, const float (*Minv)[ND]) {
  constexpr int NBA = Env<NBLK>::NBA;
  NcRows nc;
  nc_setup(nc, e.q, e.qd, e.mt, e.mi, Minv);
  // ---- contact rows: one normal + two tangents per cached manifold point ----------------------
  ContactRows<NBLK> cr;
  cr.n = 0; cr.nrob = 0;
  constexpr int NP = num_pairs(NBLK);
#pragma unroll 1
  for (int k = 0; k < NP; k++) {
    int n = __float_as_int(e.mw(k, 0));
    if (n == 0) continue;
    PairInfo pi = pair_info<NBLK>(k);
    V3 pa, pb; M3 Ra, Rb;
    geom_pose(e, f, pi.ka, pi.ia, pa, Ra);
    geom_pose(e, f, pi.kb, pi.ib, pb, Rb);
    const V3 pa_anchor = pa + geom_anchor(pi.ka), pb_anchor = pb + geom_anchor(pi.kb);  // what lA / lB refer to
    const float mu = geom_friction(pi.ka) * geom_friction(pi.kb);
    const bool robotA = geom_robot(pi.ka);  // the gripper base moves with the 7 arm joints only (both finger columns zero)
    const bool blockA = pi.ka == G_BLOCK, blockB = pi.kb == G_BLOCK;
#pragma unroll 1
    for (int i = 0; i < n; i++) {
      if (cr.n >= ContactRows<NBLK>::MP || (robotA && cr.nrob >= ContactRows<NBLK>::MR)) { e.overflow++; continue; }
      const int c = cr.n++;
      const int o = 1 + 10 * i;
      V3 lA = v3(e.mw(k, o), e.mw(k, o + 1), e.mw(k, o + 2)), lB = v3(e.mw(k, o + 3), e.mw(k, o + 4), e.mw(k, o + 5));
      V3 nB = v3(e.mw(k, o + 6), e.mw(k, o + 7), e.mw(k, o + 8));
      const float dist = e.mw(k, o + 9);
      V3 wa = mul(Ra, lA) + pa_anchor, wb = mul(Rb, lB) + pb_anchor;
      V3 dirs[3];
      dirs[0] = nB;
      plane_space(nB, dirs[1], dirs[2]);
      V3 rA = wa - pa, rB = wb - pb;
      int ri = -1;
      V3 Jp[ND];  // velocity of the contact point per unit joint velocity
      if (robotA) {
        ri = cr.nrob++;
#pragma unroll
        for (int j = 0; j < 7; j++) Jp[j] = cross(f.a[j], wa - f.p[j]);
        Jp[7] = pi.ka == G_FINGER1 ? f.a[PMG_BODY_FINGER1] : v3(0, 0, 0);
        Jp[8] = pi.ka == G_FINGER2 ? f.a[PMG_BODY_FINGER2] : v3(0, 0, 0);
      }
      const int ends = (ri & 0xff) | (((blockA ? pi.ia : -1) & 0xff) << 8) | (((blockB ? pi.ib : -1) & 0xff) << 16);
#pragma unroll 1
      for (int kk = 0; kk < 3; kk++) {
        V3 d = dirs[kk];
        RowRec r;
        r.ends = ends; r.mu = mu; r.pad = 0;
        r.dx = d.x; r.dy = d.y; r.dz = d.z;
        V3 xa = blockA ? cross(rA, d) : v3(0, 0, 0), xb = blockB ? cross(rB, d) : v3(0, 0, 0);
        r.ax = xa.x; r.ay = xa.y; r.az = xa.z; r.bx = xb.x; r.by = xb.y; r.bz = xb.z;
        float denom = 0.0f;
        if (ri >= 0) {
          RobotRow& rr = cr.rrow[ri][kk];
          float J[ND];
#pragma unroll
          for (int j = 0; j < ND; j++) { J[j] = dot(d, Jp[j]); rr.J[j] = J[j]; }
#pragma unroll
          for (int rr_ = 0; rr_ < ND; rr_++) {
            float s = 0.0f;
#pragma unroll
            for (int j = 0; j < ND; j++) s += Minv[rr_][j] * J[j];
            rr.MJ[rr_] = s;
            denom += J[rr_] * s;
          }
        }
        if (blockA) denom += BLOCK_INV_MASS + BLOCK_INV_INERTIA * dot(xa, xa);
        if (blockB) denom += BLOCK_INV_MASS + BLOCK_INV_INERTIA * dot(xb, xb);
        r.denom = denom;
        r.dinv = 1.0f / denom;
        r.app = 0.0f;
        r.rhs = 0.0f;
        cr.row[c][kk] = r;
        float rel_vel = row_velocity(cr, c, kk, r, e.qd, e.bv, e.bw);
        if (kk == 0) {
          float pen = dist + LINEAR_SLOP;
          float pos_err = 0.0f, vel_err = -rel_vel;
          if (pen > 0.0f) vel_err -= pen * INV_DT; else pos_err = -pen * CONTACT_ERP * INV_DT;
          cr.row[c][0].rhs = (pos_err + vel_err) * r.dinv;
        } else cr.row[c][kk].rhs = -rel_vel * r.dinv;
      }
    }
  }
  // ---- PGS: <= 5 iterations, early exit on the largest squared velocity change ---------------
  float dqd[ND];
  V3 dlin[NBA], dang[NBA];
#pragma unroll
  for (int j = 0; j < ND; j++) dqd[j] = 0.0f;
#pragma unroll
  for (int b = 0; b < NBA; b++) { dlin[b] = v3(0, 0, 0); dang[b] = v3(0, 0, 0); }
#pragma unroll 1
  for (int it = 0; it < SOLVER_ITERS; it++) {
    float res = 0.0f;
    nc_sweep(nc, (it & 1) != 0, Minv, dqd, res);  // backwards on even iterations, forwards on odd
#pragma unroll 1
    for (int c = 0; c < cr.n; c++) {
      RowRec r = cr.row[c][0];
      float dl = r.rhs - row_velocity(cr, c, 0, r, dqd, dlin, dang) * r.dinv;
      float sum = fminf(fmaxf(r.app + dl, 0.0f), 1e10f);
      dl = sum - r.app;
      cr.row[c][0].app = sum;
      row_apply(cr, c, 0, r, dl, dqd, dlin, dang);
      float rr = dl * r.denom;
      res = fmaxf(res, rr * rr);
    }
#pragma unroll 1
    for (int c = 0; c < cr.n; c++) {  // implicit friction cone: both tangent rows of a point together
      const float total = cr.row[c][0].app;
      if (!(total > 0.0f)) continue;
      RowRec ra = cr.row[c][1], rb = cr.row[c][2];
      const float lim = ra.mu * total;
      float dA = ra.rhs - row_velocity(cr, c, 1, ra, dqd, dlin, dang) * ra.dinv;
      float dB = rb.rhs - row_velocity(cr, c, 2, rb, dqd, dlin, dang) * rb.dinv;
      float sA = ra.app + dA, sB = rb.app + dB;
      const float s2 = sA * sA + sB * sB;
      if (s2 >= lim * lim) {
        // |lim sin(atan2(sA, sB))| = lim |sA| / sqrt(sA^2 + sB^2), likewise the cosine for sB
        const float sc = s2 > 0.0f ? lim * rsqrtf(s2) : 0.0f;
        const float cA = fabsf(sA) * sc, cB = s2 > 0.0f ? fabsf(sB) * sc : lim;
        sA = fminf(fmaxf(sA, -cA), cA);
        sB = fminf(fmaxf(sB, -cB), cB);
        dA = sA - ra.app; dB = sB - rb.app;
      }
      cr.row[c][1].app = sA; cr.row[c][2].app = sB;
      row_apply(cr, c, 1, ra, dA, dqd, dlin, dang);
      row_apply(cr, c, 2, rb, dB, dqd, dlin, dang);
      float r1 = dA * ra.denom, r2 = dB * rb.denom;
      res = fmaxf(res, fmaxf(r1 * r1, r2 * r2));
    }
    if (res <= RESIDUAL_THRESHOLD) break;
  }
#pragma unroll
  for (int j = 0; j < ND; j++) e.qd[j] = fminf(fmaxf(e.qd[j] + dqd[j], -MAX_COORD_VEL), MAX_COORD_VEL);
#pragma unroll
  for (int b = 0; b < NBLK; b++) { e.bv[b] += dlin[b]; e.bw[b] += dang[b]; }
}

// ---- one 2 ms substep (btMultiBodyDynamicsWorld::internalSingleStepSimulation) ---------------
template <int NBLK>
__device__ void substep(Env<NBLK>& e) {
  Frames f;
  float Minv[ND][ND];
  {
    float bias[ND], M[ND][ND];
    robot_dynamics(e.q, e.qd, f, bias, M);  // link frames first: the collision pass needs them
    collide(e, f);
    invert_spd9(M, Minv);
#pragma unroll
    for (int i = 0; i < ND; i++) {
      float s = 0.0f;
#pragma unroll
      for (int j = 0; j < ND; j++) s += Minv[i][j] * (e.dtau[j] - bias[j]);
      e.qd[i] = fminf(fmaxf(e.qd[i] + s * DT, -MAX_COORD_VEL), MAX_COORD_VEL);
    }
  }
#pragma unroll
  for (int b = 0; b < NBLK; b++) {  // free cubes: gravity + Bullet's velocity damping; isotropic inertia => no gyro term
    float kl = LINK_DAMPING + LINK_DAMPING * norm(e.bv[b]), ka = LINK_DAMPING + LINK_DAMPING * norm(e.bw[b]);
    e.bv[b] += DT * (v3(0, 0, -GRAVITY) - kl * e.bv[b]);
    e.bw[b] -= (DT * ka) * e.bw[b];
  }
  solve_constraints(e, f, Minv);
#pragma unroll
  for (int j = 0; j < ND; j++) e.q[j] += e.qd[j] * DT;
#pragma unroll
  for (int b = 0; b < NBLK; b++) {
    e.bpos[b] += DT * e.bv[b];
    float w = norm(e.bw[b]);
    if (w * DT > 0.25f * PI_F) w = 0.25f * PI_F / DT;  // ANGULAR_MOTION_THRESHOLD
    float s = w < 0.001f ? 0.5f * DT - DT * DT * DT * 0.020833333333f * w * w : sinf(0.5f * w * DT) / w;
    float ax = e.bw[b].x * s, ay = e.bw[b].y * s, az = e.bw[b].z * s, aw = cosf(0.5f * w * DT);
    float bx = e.bquat[b][0], by = e.bquat[b][1], bz = e.bquat[b][2], bw_ = e.bquat[b][3];
    float nx = aw * bx + ax * bw_ + ay * bz - az * by;
    float ny = aw * by + ay * bw_ + az * bx - ax * bz;
    float nz = aw * bz + az * bw_ + ax * by - ay * bx;
    float nw = aw * bw_ - ax * bx - ay * by - az * bz;
    float inv = 1.0f / sqrtf(nx * nx + ny * ny + nz * nz + nw * nw);
    e.bquat[b][0] = nx * inv; e.bquat[b][1] = ny * inv; e.bquat[b][2] = nz * inv; e.bquat[b][3] = nw * inv;
  }
}

// ---- kernel I/O (shared by the thread-per-env and the lane-cooperative kernels) -----------------
struct StepIO {
  float* state; float* manifold; int batch; int state_words;
  // Layout of the two persistent arrays: tiles of `tile` consecutive environments, [tile][word][env in tile], so that a
  // warp's accesses to one word of its environments are one contiguous run: tile = 4 for the lane-cooperative kernels
  // (lane l of each of the warp's four octets reads word w0 + l: 32 consecutive floats, one 128-byte line per
  // instruction), 32 for the thread-per-env kernels; tile = batch is the plain [word][env] array of round 1
  // (PMG_STATE_TILE=0), which the cooperative kernels read 16 bytes per 32-byte sector.
  int tile; int man_words;
  const float* action; float* obs; float* reward; uint8_t* done; uint8_t* success;
  float thr; int binary; int max_steps; int* overflow;
  int epw;  // environments per warp: lanes [0, epw) of every warp own one environment each
  int epb;  // lane-cooperative kernels: environments (octets) per one-warp block, 4 / 2 / 1 (see coop_geometry)
  int bulk;         // 1: stage the state tile with TMA bulk copies (full warps only)
  int tile_offset;  // float offset of the state tile inside dynamic shared memory
  // run-time variants of a task (the block-stack kernels also serve block_rearrange):
  int grasp;        // the last action column drives the jaws and finger_closeness / finger_vel are observed
  int jc;           // joint-space control: 7 joint deltas (+ grip), joint poses prepended to observation / policy_state
  int grip_goal;    // grip-informed goal: achieved / desired goal gain gripper xyz + finger closeness
  int td;           // task decomposition / curriculum: desired goal = the sub-goal named by the state word behind the goal (1); 2: BlockRearrange curriculum, that word is the bit mask of the blocks with targets
  int cur;          // curriculum: that word is set by the reset (from the spawn row) instead of -1
  int adim, goal_dim, row_width;  // action columns, goal length, packed row width (all after the variants above)
  float* row_spill;               // cooperative block kernel: contact-row records beyond the shared-memory ones
  // ---- multi-GPU gather over peer memory (pmg_step_gather; include/pmg.h) ----
  // obs / reward / done / success above point at this rank's slice of its OWN gather buffer; g_* are the same slices
  // of the g_n peers' buffers (peer-mapped device pointers: plain stores travel over NVLink).
  int g_n;               // number of remote destinations, 0 = no gather
  int g_in_step;         // 1: the step kernel pushes and publishes; 0: the auto-reset pass behind it does
  int g_world;           // ranks (flags to publish / wait for)
  unsigned g_seq;        // sequence number of this step
  float* g_obs[7]; float* g_reward[7]; uint8_t* g_done[7]; uint8_t* g_success[7];
  unsigned* g_flag[8];   // this rank's flag word in every rank's buffer (own buffer included)
  const unsigned* g_flags_local;  // the world flag words of the own buffer
  unsigned* g_counter;   // arrivals of this launch (one per environment), reset by the last one
  int* g_err;            // mapped host word: set when a peer's flag did not arrive in time
  int* bad_action;       // mapped host word: set when an action entry lies outside [-1, 1] (NaN included); see pmg_action_error
};

// offset of environment `env`'s word 0 in the state / manifold array; its consecutive words are io.tile floats apart
__device__ __forceinline__ size_t state_off(const StepIO& io, int env) { return (size_t)(env / io.tile) * io.state_words * io.tile + env % io.tile; }
__device__ __forceinline__ size_t man_off(const StepIO& io, int env) { return (size_t)(env / io.tile) * io.man_words * io.tile + env % io.tile; }

// Box(-1, 1).contains of one environment's action row (kuka.py:168 asserts it), by the `nl` lanes that own the environment
__device__ __forceinline__ void check_action_row(const StepIO& io, int env, int lane, int nl) {
  const float* a = io.action + (size_t)env * io.adim;
  bool bad = false;
  for (int k = lane; k < io.adim; k += nl) bad = bad || !(a[k] >= -1.0f && a[k] <= 1.0f);
  if (bad) *(volatile int*)io.bad_action = 1;
}

#ifndef PMG_EMULATE
// Called once per environment after its row, reward and flags are stored locally AND on the peers.  The last arrival
// of the launch publishes this rank's sequence number to every rank and then waits for theirs, so that the kernel's
// completion means "the local gather buffer holds the whole global batch of this step".  Peers never wait for this
// kernel to finish, only for the flag it has already published, so there is no cyclic wait.  The data buffers are
// double-buffered by the parity of g_seq (a peer can be at most one step ahead).
__device__ __forceinline__ void gather_arrive(const StepIO& io) {
  __threadfence_system();  // this environment's peer stores are ordered before the arrival
  const unsigned old = atomicAdd(io.g_counter, 1u);
  if (old != (unsigned)io.batch - 1u) return;
  atomicExch(io.g_counter, 0u);
  __threadfence_system();  // every environment's stores (observed through the counter) before the flags
  for (int d = 0; d < io.g_world; d++) *(volatile unsigned*)io.g_flag[d] = io.g_seq;
  unsigned long long t0, t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (int src = 0; src < io.g_world; src++) {
    const volatile unsigned* f = io.g_flags_local + src;
    while ((int)(*f - io.g_seq) < 0) {
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > 5000000000ull) { *(volatile int*)io.g_err = 1 + src; return; }  // 5 s: a peer died; report, do not hang the GPU
      __nanosleep(200);
    }
  }
  __threadfence_system();
}
#else
__device__ __forceinline__ void gather_arrive(const StepIO&) {}
#endif

template <int TASK, int NBLK> struct Dims {
  static constexpr int O = TASK == 0 ? 3 : (TASK == 3 ? 8 + 16 * NBLK : 20);
  static constexpr int P = TASK == 0 ? 3 : (TASK == 3 ? 4 + 3 * NBLK : 7);
  static constexpr int G = TASK == 3 ? 3 * NBLK : 3;
  static constexpr int W = O + P + 2 * G;
  static constexpr int A = TASK >= 2 ? 4 : 3;
  static constexpr int STATE = ST_BLK + 13 * NBLK + G + 1;
};


}  // namespace pmg
