// pmg_physics.cuh -- per-environment device routines of the batched Kuka simulator (sm_100a).
//
// One CUDA thread owns one environment for the whole env.step(): the hot loop (100 substeps
// of forward dynamics + contact generation + 5-iteration projected Gauss-Seidel) is strictly
// sequential per environment, so lanes are spent on independent environments, not on one.
// Persistent state is struct-of-arrays in HBM ([word][env]) so that a warp's 32 environments
// read and write 128-byte lines; the working set of a substep lives in registers / L1-resident
// local memory.  No tensor cores: the largest matrix on this path is 9x9.
//
// What is computed follows the reference's call sequence (paths relative to
// /root/reference/pybullet_multigoal_gym/): robots/kuka.py:167-225 (action map, IK, motors,
// 5 x stepSimulation), envs/base_envs/base_env.py:215-219 (0.002 s x 20 substeps, 5 solver
// iterations, contact ERP 0.9), and the Bullet behaviours listed in DESIGN.md.  The arithmetic
// is organised differently from Bullet (composite-rigid-body mass matrix + Cholesky instead of
// per-row articulated-body impulse responses) but is mathematically the same system.
//
// Code-shape rules that come from measurements (profiles/): the L1.5 instruction cache holds 32 KB
// (2 K instructions) and with ~2 warps per SM nothing hides a fetch miss, so per-body loops are
// rolled (all threads of a warp are at the same body => uniform branches) and only small hot
// blocks (the 9 motor rows, the 9x9 Cholesky) are unrolled in registers.
#pragma once

#ifdef PMG_EMULATE
#include "pmg_emu_shim.h"  // host build of the same device code (tests/emu): CUDA keywords and intrinsics as plain C++
#else
#include <cuda_runtime.h>
#endif
#include <float.h>
#include <math.h>
#include <stdint.h>

#include "../../include/pmg_model_constants.h"

namespace pmg {

constexpr int NB = PMG_NBODY;  // 10 robot bodies
constexpr int ND = PMG_NDOF;   // 9 dofs

// ---- physics parameters (base_env.py:215-219, kuka.py:223-225,282-301; Bullet defaults) ----
constexpr float DT = 0.002f;
constexpr float INV_DT = 500.0f;
constexpr float OUTER_DT = 0.04f;
constexpr int SUBSTEPS_PER_CALL = 20;
constexpr int CALLS_PER_ENV_STEP = 5;
constexpr int SOLVER_ITERS = 5;
constexpr float CONTACT_ERP = 0.9f;
constexpr float LINEAR_SLOP = 1e-5f;
constexpr float RESIDUAL_THRESHOLD = 1e-7f;
constexpr float GRAVITY = 9.81f;
constexpr float LINK_DAMPING = 0.04f;
constexpr float MAX_COORD_VEL = 100.0f;
constexpr float LIMIT_MAX_IMPULSE = 100.0f;
constexpr float SPLIT_IMPULSE_PEN_THRESHOLD = -0.04f;
constexpr float MOTOR_KP = 0.03f;
constexpr float MOTOR_KD = 1.0f;
constexpr float ARM_FORCE = 200.0f;
constexpr float FINGER_FORCE = 50.0f;
constexpr float BREAKING_THRESHOLD_FACTOR = 0.02f;
constexpr float BROADPHASE_MARGIN = 0.02f;
constexpr float IK_JOINT_DAMPING = 0.5f;
constexpr float IK_MAX_STEP = 0.78539816339744831f;
constexpr float BOX_FUDGE = 1.05f;
constexpr float PI_F = 3.14159265358979323846f;
constexpr float GRIPPER_ABS_LIMIT = 0.035f;
constexpr float BLOCK_SPAWN_Z = 0.175f;
constexpr float BLOCK_HALF = (float)PMG_BLOCK_HALF;
constexpr float BLOCK_INV_MASS = (float)(1.0 / PMG_BLOCK_MASS);
constexpr float BLOCK_INV_INERTIA = (float)(1.0 / PMG_BLOCK_INERTIA);

// ---- model tables (constant memory; the index is uniform across the warp => broadcast loads) --
__constant__ float c_jxyz[NB][3] = PMG_BODY_JXYZ;
__constant__ float c_jrot[NB][9] = PMG_BODY_JROT;
__constant__ float c_mass[NB] = PMG_BODY_MASS;
__constant__ float c_com[NB][3] = PMG_BODY_COM;
__constant__ float c_inertia[NB][3] = PMG_BODY_INERTIA;
__constant__ float c_dof_lower[ND] = PMG_DOF_LOWER;
__constant__ float c_dof_upper[ND] = PMG_DOF_UPPER;
__constant__ float c_dof_damping[ND] = PMG_DOF_DAMPING;
__constant__ int c_nc_order[2 * ND] = PMG_NONCONTACT_ORDER;
__constant__ float c_rest_pose0[7] = {0.f, -0.5592432f, 0.f, 1.733180f, 0.f, -0.8501557f, 0.f};  // kuka.py:27

__host__ __device__ constexpr int body_parent(int b) { return b == 0 ? -1 : (b <= 7 ? b - 1 : 7); }
__host__ __device__ constexpr int body_jtype(int b) { return b <= 6 ? 0 : (b == 7 ? 2 : 1); }
__host__ __device__ constexpr int body_dof(int b) { return b <= 6 ? b : (b == 7 ? -1 : b - 1); }
__host__ __device__ constexpr int dof_body(int d) { return d <= 6 ? d : d + 1; }

// ---- small vector types ---------------------------------------------------------------------
struct V3 { float x, y, z; };
__device__ __forceinline__ V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ V3 operator*(float s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return v3(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ V3& operator+=(V3& a, V3 b) { a.x += b.x; a.y += b.y; a.z += b.z; return a; }
__device__ __forceinline__ V3& operator-=(V3& a, V3 b) { a.x -= b.x; a.y -= b.y; a.z -= b.z; return a; }
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ float norm(V3 a) { return sqrtf(dot(a, a)); }
__device__ __forceinline__ float comp(V3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

struct M3 { V3 r0, r1, r2; };  // rows
__device__ __forceinline__ M3 m3_identity() { M3 m; m.r0 = v3(1, 0, 0); m.r1 = v3(0, 1, 0); m.r2 = v3(0, 0, 1); return m; }
__device__ __forceinline__ V3 mul(const M3& m, V3 v) { return v3(dot(m.r0, v), dot(m.r1, v), dot(m.r2, v)); }
__device__ __forceinline__ V3 mulT(const M3& m, V3 v) { return v.x * m.r0 + v.y * m.r1 + v.z * m.r2; }
__device__ __forceinline__ V3 col(const M3& m, int i) { return v3(comp(m.r0, i), comp(m.r1, i), comp(m.r2, i)); }
__device__ __forceinline__ M3 mul(const M3& a, const M3& b) {
  M3 r;
  r.r0 = a.r0.x * b.r0 + a.r0.y * b.r1 + a.r0.z * b.r2;
  r.r1 = a.r1.x * b.r0 + a.r1.y * b.r1 + a.r1.z * b.r2;
  r.r2 = a.r2.x * b.r0 + a.r2.y * b.r1 + a.r2.z * b.r2;
  return r;
}
__device__ __forceinline__ M3 quat_to_m3(float x, float y, float z, float w) {
  float d = x * x + y * y + z * z + w * w, s = 2.0f / d;
  float xs = x * s, ys = y * s, zs = z * s;
  float wx = w * xs, wy = w * ys, wz = w * zs, xx = x * xs, xy = x * ys, xz = x * zs, yy = y * ys, yz = y * zs, zz = z * zs;
  M3 m;
  m.r0 = v3(1 - (yy + zz), xy - wz, xz + wy);
  m.r1 = v3(xy + wz, 1 - (xx + zz), yz - wx);
  m.r2 = v3(xz - wy, yz + wx, 1 - (xx + yy));
  return m;
}
__device__ __forceinline__ void m3_to_quat(const M3& m, float q[4]) {  // xyzw
  float e[3][3] = {{m.r0.x, m.r0.y, m.r0.z}, {m.r1.x, m.r1.y, m.r1.z}, {m.r2.x, m.r2.y, m.r2.z}};
  float trace = e[0][0] + e[1][1] + e[2][2];
  if (trace > 0) {
    float s = sqrtf(trace + 1.0f);
    q[3] = s * 0.5f; s = 0.5f / s;
    q[0] = (e[2][1] - e[1][2]) * s; q[1] = (e[0][2] - e[2][0]) * s; q[2] = (e[1][0] - e[0][1]) * s;
  } else {
    int i = e[0][0] < e[1][1] ? (e[1][1] < e[2][2] ? 2 : 1) : (e[0][0] < e[2][2] ? 2 : 0);
    int j = (i + 1) % 3, k = (i + 2) % 3;
    float s = sqrtf(e[i][i] - e[j][j] - e[k][k] + 1.0f);
    q[i] = s * 0.5f; s = 0.5f / s;
    q[3] = (e[k][j] - e[j][k]) * s;
    q[j] = (e[j][i] + e[i][j]) * s;
    q[k] = (e[k][i] + e[i][k]) * s;
  }
}

// ---- kinematics -----------------------------------------------------------------------------
struct Frames {
  M3 R[NB];   // link frame orientation (world)
  V3 p[NB];   // link frame origin (world)
  V3 a[NB];   // joint axis (world); zero for the fixed gripper base
};

__device__ __forceinline__ M3 load_jrot(int b) {
  M3 m;
  m.r0 = v3(c_jrot[b][0], c_jrot[b][1], c_jrot[b][2]);
  m.r1 = v3(c_jrot[b][3], c_jrot[b][4], c_jrot[b][5]);
  m.r2 = v3(c_jrot[b][6], c_jrot[b][7], c_jrot[b][8]);
  return m;
}

// Link frames of bodies [0, NBODIES).  Arm joints rotate about their local z, the fingers slide
// along -/+ y of the gripper base (iiwa14_parallel_jaw.urdf:94-288,418-456).
template <int NBODIES>
__device__ __noinline__ void forward_kinematics(const float* q, Frames& f) {
#pragma unroll 1
  for (int b = 0; b < NBODIES; b++) {
    const int par = body_parent(b);
    M3 Rp = par < 0 ? m3_identity() : f.R[par];
    V3 pp = par < 0 ? v3(0, 0, 0) : f.p[par];
    M3 Rj = mul(Rp, load_jrot(b));
    V3 pj = pp + mul(Rp, v3(c_jxyz[b][0], c_jxyz[b][1], c_jxyz[b][2]));
    if (body_jtype(b) == 0) {
      float s, c;
      sincosf(q[body_dof(b)], &s, &c);
      // Rj * Rz(q): new x column = c*x + s*y, new y column = -s*x + c*y
      M3 R;
      R.r0 = v3(c * Rj.r0.x + s * Rj.r0.y, -s * Rj.r0.x + c * Rj.r0.y, Rj.r0.z);
      R.r1 = v3(c * Rj.r1.x + s * Rj.r1.y, -s * Rj.r1.x + c * Rj.r1.y, Rj.r1.z);
      R.r2 = v3(c * Rj.r2.x + s * Rj.r2.y, -s * Rj.r2.x + c * Rj.r2.y, Rj.r2.z);
      f.R[b] = R; f.p[b] = pj; f.a[b] = col(Rj, 2);
    } else if (body_jtype(b) == 1) {
      V3 ay = col(Rj, 1);
      V3 ax = b == PMG_BODY_FINGER1 ? -ay : ay;  // axis (0,-1,0) / (0,+1,0)
      f.R[b] = Rj; f.a[b] = ax; f.p[b] = pj + q[body_dof(b)] * ax;
    } else {
      f.R[b] = Rj; f.p[b] = pj; f.a[b] = v3(0, 0, 0);
    }
  }
}

__device__ __forceinline__ V3 tip_position(const Frames& f) {
  const float t[3] = PMG_TIP_OFFSET;
  return f.p[PMG_BODY_LINK7] + mul(f.R[PMG_BODY_LINK7], v3(t[0], t[1], t[2]));
}

// velocity of world point `pt` rigidly attached to `body`, and the body's angular velocity
template <int BODY>
__device__ __forceinline__ void point_velocity(const Frames& f, const float* qd, V3 pt, V3& lin, V3& ang) {
  lin = v3(0, 0, 0); ang = v3(0, 0, 0);
#pragma unroll
  for (int j = 0; j <= 6; j++) {
    lin += qd[j] * cross(f.a[j], pt - f.p[j]);
    ang += qd[j] * f.a[j];
  }
  if (BODY == PMG_BODY_FINGER1 || BODY == PMG_BODY_FINGER2) lin += qd[body_dof(BODY)] * f.a[BODY];
}

// ---- inverse kinematics (pybullet calculateInverseKinematics, DLS, no null space) -----------
// kuka.py:266-279: <= 40 iterations, stop when the tip is within 1e-5 of the target; per
// iteration dtheta = (J^T J + 0.5 I)^-1 J^T e with e = [position error; orientation error as
// axis*angle], largest |dtheta| clamped to 45 degrees.  Finger columns of J are zero, so the
// 9x9 system splits into the 7x7 arm block solved here and dtheta_finger = 0.
__device__ void inverse_kinematics(float* q /* in: seed, out: result (first 7 used) */, V3 target, const float tq[4]) {
  float diff = 1e30f;
  for (int it = 0; it < 40 && diff > 1e-5f; it++) {
    Frames f;
    forward_kinematics<7>(q, f);
    V3 tip = tip_position(f);
    V3 ep = target - tip;
    diff = norm(ep);
    float qc[4];
    m3_to_quat(f.R[PMG_BODY_LINK7], qc);
    // dq = target * conj(current)
    float ax = -qc[0], ay = -qc[1], az = -qc[2], aw = qc[3];
    float dx = tq[3] * ax + tq[0] * aw + tq[1] * az - tq[2] * ay;
    float dy = tq[3] * ay + tq[1] * aw + tq[2] * ax - tq[0] * az;
    float dz = tq[3] * az + tq[2] * aw + tq[0] * ay - tq[1] * ax;
    float dw = tq[3] * aw - tq[0] * ax - tq[1] * ay - tq[2] * az;
    // axis*angle; 2*atan2(|v|, w) equals Bullet's 2*acos(w) but stays accurate in fp32 near 0
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
      float s = ang / vn;
      er = v3(dx * s, dy * s, dz * s);
    }
    // J columns: linear a_j x (tip - o_j), angular a_j
    V3 Jl[7], Ja[7];
#pragma unroll
    for (int j = 0; j < 7; j++) { Jl[j] = cross(f.a[j], tip - f.p[j]); Ja[j] = f.a[j]; }
    float A[7][7], rhs[7];
#pragma unroll
    for (int i = 0; i < 7; i++) {
#pragma unroll
      for (int j = 0; j <= i; j++) A[i][j] = dot(Jl[i], Jl[j]) + dot(Ja[i], Ja[j]) + (i == j ? IK_JOINT_DAMPING : 0.0f);
      rhs[i] = dot(Jl[i], ep) + dot(Ja[i], er);
    }
    // Cholesky solve of the SPD 7x7 system
#pragma unroll
    for (int j = 0; j < 7; j++) {
      float d = A[j][j];
#pragma unroll
      for (int k = 0; k < j; k++) d -= A[j][k] * A[j][k];
      d = sqrtf(d);
      A[j][j] = d;
      float inv = 1.0f / d;
#pragma unroll
      for (int i = j + 1; i < 7; i++) {
        float s = A[i][j];
#pragma unroll
        for (int k = 0; k < j; k++) s -= A[i][k] * A[j][k];
        A[i][j] = s * inv;
      }
    }
#pragma unroll
    for (int i = 0; i < 7; i++) {
      float s = rhs[i];
#pragma unroll
      for (int k = 0; k < i; k++) s -= A[i][k] * rhs[k];
      rhs[i] = s / A[i][i];
    }
#pragma unroll
    for (int i = 6; i >= 0; i--) {
      float s = rhs[i];
#pragma unroll
      for (int k = i + 1; k < 7; k++) s -= A[k][i] * rhs[k];
      rhs[i] = s / A[i][i];
    }
    float mx = 0.0f;
#pragma unroll
    for (int i = 0; i < 7; i++) mx = fmaxf(mx, fabsf(rhs[i]));
    float scale = mx > IK_MAX_STEP ? IK_MAX_STEP / mx : 1.0f;
#pragma unroll
    for (int i = 0; i < 7; i++) q[i] += scale * rhs[i];
  }
}

// ---- fused robot dynamics: two sweeps over the kinematic tree ---------------------------------
// Outward sweep (base -> fingers): link frames, world inertia of every body and the Newton-Euler
// velocity / acceleration recursion with qdd = 0, with the parent's state carried in registers.
// Inward sweep (fingers -> base): accumulates subtree force/moment (-> bias forces) and subtree
// composite inertia (-> one mass-matrix row per joint) with the same origin shifts.
// Loops are rolled on purpose: every thread of the warp is at the same body, so the type branches
// are uniform, and the step kernel is bound by instruction fetch + dependent-issue latency.
struct Sym3 { float xx, xy, xz, yy, yz, zz; };
__device__ __forceinline__ V3 mul(const Sym3& m, V3 v) {
  return v3(m.xx * v.x + m.xy * v.y + m.xz * v.z, m.xy * v.x + m.yy * v.y + m.yz * v.z, m.xz * v.x + m.yz * v.y + m.zz * v.z);
}

__device__ __noinline__ void robot_dynamics(const float* q, const float* qd, Frames& f, float* bias, float M[ND][ND]) {
  V3 Fb[NB], Nb[NB], rcb[NB];   // net force, net moment about the link origin, COM offset (world axes)
  Sym3 Iwb[NB];                 // world inertia about the COM
  {
    M3 Rp = m3_identity();
    V3 pp = v3(0, 0, 0), wp = v3(0, 0, 0), alp = v3(0, 0, 0), accp = v3(0, 0, GRAVITY), velp = v3(0, 0, 0);
#pragma unroll 1
    for (int b = 0; b < NB; b++) {
      const int jt = body_jtype(b), dof = body_dof(b);
      M3 Rj = mul(Rp, load_jrot(b));
      V3 r = mul(Rp, v3(c_jxyz[b][0], c_jxyz[b][1], c_jxyz[b][2]));  // parent origin -> joint origin
      M3 R = Rj;
      V3 ax = v3(0, 0, 0), aq = v3(0, 0, 0);
      if (jt == 0) {
        float sn, cs;
        sincosf(q[dof], &sn, &cs);
        R.r0 = v3(cs * Rj.r0.x + sn * Rj.r0.y, -sn * Rj.r0.x + cs * Rj.r0.y, Rj.r0.z);
        R.r1 = v3(cs * Rj.r1.x + sn * Rj.r1.y, -sn * Rj.r1.x + cs * Rj.r1.y, Rj.r1.z);
        R.r2 = v3(cs * Rj.r2.x + sn * Rj.r2.y, -sn * Rj.r2.x + cs * Rj.r2.y, Rj.r2.z);
        ax = col(Rj, 2);
        aq = qd[dof] * ax;
      } else if (jt == 1) {
        V3 ay = col(Rj, 1);
        ax = b == PMG_BODY_FINGER1 ? -ay : ay;
        r += q[dof] * ax;
        aq = qd[dof] * ax;
      }
      V3 p = pp + r;
      V3 wxr = cross(wp, r);
      V3 a_o = accp + cross(alp, r) + cross(wp, wxr);
      V3 v_o = velp + wxr;
      V3 w = wp, al = alp;
      if (jt == 0) { w = wp + aq; al = alp + cross(wp, aq); }
      else if (jt == 1) { a_o += 2.0f * cross(wp, aq); v_o += aq; }
      // world inertia R diag(I) R^T and COM offset
      V3 rc = mul(R, v3(c_com[b][0], c_com[b][1], c_com[b][2]));
      const float i0 = c_inertia[b][0], i1 = c_inertia[b][1], i2 = c_inertia[b][2];
      V3 s0 = v3(R.r0.x * i0, R.r0.y * i1, R.r0.z * i2), s1 = v3(R.r1.x * i0, R.r1.y * i1, R.r1.z * i2), s2 = v3(R.r2.x * i0, R.r2.y * i1, R.r2.z * i2);
      Sym3 Iw;
      Iw.xx = dot(s0, R.r0); Iw.xy = dot(s0, R.r1); Iw.xz = dot(s0, R.r2);
      Iw.yy = dot(s1, R.r1); Iw.yz = dot(s1, R.r2); Iw.zz = dot(s2, R.r2);
      // Newton-Euler at the COM, Bullet's per-link velocity damping as an external force
      V3 wxrc = cross(w, rc);
      V3 a_c = a_o + cross(al, rc) + cross(w, wxrc);
      V3 v_c = v_o + wxrc;
      const float m = c_mass[b];
      const float kl = LINK_DAMPING + LINK_DAMPING * norm(v_c), ka = LINK_DAMPING + LINK_DAMPING * norm(w);
      V3 Fc = m * a_c + (m * kl) * v_c;
      V3 Iww = mul(Iw, w);
      V3 Nc = mul(Iw, al) + cross(w, Iww) + ka * Iww;
      f.R[b] = R; f.p[b] = p; f.a[b] = ax;
      Fb[b] = Fc; Nb[b] = Nc + cross(rc, Fc); rcb[b] = rc; Iwb[b] = Iw;
      if (b <= PMG_BODY_GBASE) { Rp = R; pp = p; wp = w; alp = al; accp = a_o; velp = v_o; }  // both fingers hang off the gripper base
    }
  }
  M[8][7] = 0.0f;  // the two fingers are siblings
  // inbox: quantities of already-visited children, expressed about the origin of their parent
  V3 inF = v3(0, 0, 0), inN = v3(0, 0, 0), inh = v3(0, 0, 0);
  float inm = 0.0f;
  Sym3 inI; inI.xx = inI.xy = inI.xz = inI.yy = inI.yz = inI.zz = 0.0f;
#pragma unroll 1
  for (int b = NB - 1; b >= 0; b--) {
    const int jt = body_jtype(b), dof = body_dof(b);
    const bool leaf = b >= PMG_BODY_FINGER1;
    const float m = c_mass[b];
    V3 rc = rcb[b];
    float cc = dot(rc, rc);
    Sym3 Iw = Iwb[b];
    // own quantities about the link origin: h = m rc, I = Iw + m (rc.rc 1 - rc rc^T)
    V3 F = Fb[b], N = Nb[b], h = m * rc;
    float mt = m;
    Sym3 I;
    I.xx = Iw.xx + m * (cc - rc.x * rc.x); I.xy = Iw.xy - m * rc.x * rc.y; I.xz = Iw.xz - m * rc.x * rc.z;
    I.yy = Iw.yy + m * (cc - rc.y * rc.y); I.yz = Iw.yz - m * rc.y * rc.z; I.zz = Iw.zz + m * (cc - rc.z * rc.z);
    if (!leaf) {
      F += inF; N += inN; h += inh; mt += inm;
      I.xx += inI.xx; I.xy += inI.xy; I.xz += inI.xz; I.yy += inI.yy; I.yz += inI.yz; I.zz += inI.zz;
    }
    V3 pb = f.p[b], ab = f.a[b];
    if (jt != 2) {
      V3 n, l;  // moment about p[b] / force caused by unit acceleration of this joint
      if (jt == 0) { n = mul(I, ab); l = cross(ab, h); M[dof][dof] = dot(ab, n); bias[dof] = dot(ab, N); }
      else { l = mt * ab; n = cross(h, ab); M[dof][dof] = mt; bias[dof] = dot(ab, F); }
#pragma unroll 1
      for (int j = (b <= 6 ? b - 1 : 6); j >= 0; j--) {  // every ancestor with a dof is a revolute arm joint, dof j == body j
        V3 nj = n + cross(pb - f.p[j], l);
        M[dof][j] = dot(f.a[j], nj);
      }
    }
    if (b > 0) {
      // express the subtree about the parent's origin: r = p[b] - p[parent]
      V3 r = pb - f.p[body_parent(b)];
      V3 sN = N + cross(r, F);
      float hr = 2.0f * dot(h, r) + mt * dot(r, r);
      V3 hm = h + mt * r;
      Sym3 sI;
      sI.xx = I.xx + hr - h.x * r.x - r.x * hm.x; sI.xy = I.xy - h.x * r.y - r.x * hm.y; sI.xz = I.xz - h.x * r.z - r.x * hm.z;
      sI.yy = I.yy + hr - h.y * r.y - r.y * hm.y; sI.yz = I.yz - h.y * r.z - r.y * hm.z; sI.zz = I.zz + hr - h.z * r.z - r.z * hm.z;
      if (leaf && b == PMG_BODY_FINGER1) {  // second finger: add to what finger2 already delivered
        inF += F; inN += sN; inh += hm; inm += mt;
        inI.xx += sI.xx; inI.xy += sI.xy; inI.xz += sI.xz; inI.yy += sI.yy; inI.yz += sI.yz; inI.zz += sI.zz;
      } else { inF = F; inN = sN; inh = hm; inm = mt; inI = sI; }
    }
  }
}

// In-place Cholesky of the lower triangle, then Minv = L^-T L^-1 (full symmetric matrix out).
__device__ __noinline__ void invert_spd9(float M[ND][ND], float Minv[ND][ND]) {
#pragma unroll
  for (int j = 0; j < ND; j++) {
    float d = M[j][j];
#pragma unroll
    for (int k = 0; k < j; k++) d -= M[j][k] * M[j][k];
    d = sqrtf(d);
    float inv = 1.0f / d;
    M[j][j] = inv;  // store the reciprocal of the diagonal
#pragma unroll
    for (int i = j + 1; i < ND; i++) {
      float s = M[i][j];
#pragma unroll
      for (int k = 0; k < j; k++) s -= M[i][k] * M[j][k];
      M[i][j] = s * inv;
    }
  }
  // Linv (lower): column by column
  float Li[ND][ND];
#pragma unroll
  for (int c = 0; c < ND; c++) {
    Li[c][c] = M[c][c];
#pragma unroll
    for (int i = c + 1; i < ND; i++) {
      float s = 0.0f;
#pragma unroll
      for (int k = c; k < i; k++) s -= M[i][k] * Li[k][c];
      Li[i][c] = s * M[i][i];
    }
  }
#pragma unroll
  for (int i = 0; i < ND; i++) {
#pragma unroll
    for (int j = 0; j <= i; j++) {
      float s = 0.0f;
#pragma unroll
      for (int k = i; k < ND; k++) s += Li[k][i] * Li[k][j];
      Minv[i][j] = s; Minv[j][i] = s;
    }
  }
}

// ---- box-box contact generation (SAT + incident-face clipping, after btBoxBoxDetector) -------
struct Contact { V3 pB, nB; float dist; };

// Dynamically indexed work arrays of the narrowphase.  The thread-per-env kernels keep them in (L1-resident)
// local memory; the lane-cooperative kernel points them at shared memory, where its L1 share is tiny.
struct BoxScratch {
  float quad[8], ret[16], buf[2][16], dep[8], ang[8];
  V3 point[8];
  int idx[8];
  bool avail[8];
  Contact out[4];
};

__device__ int clip_quad_to_rect(const float h[2], const float quad[8], float out[16], float (*buf)[16]) {
  // nothing to clip (the usual case: a small face resting inside a large one): the clipper would copy the quad
  bool inside = true;
#pragma unroll
  for (int i = 0; i < 4; i++) inside = inside && fabsf(quad[2 * i]) < h[0] && fabsf(quad[2 * i + 1]) < h[1];
  if (inside) {
#pragma unroll
    for (int i = 0; i < 8; i++) out[i] = quad[i];
    return 4;
  }
  int nq = 4, nr = 0, cur = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) buf[0][i] = quad[i];
  bool full = false;
  for (int dir = 0; dir <= 1 && !full; dir++) {
    for (int sign = -1; sign <= 1 && !full; sign += 2) {
      const float* q = buf[cur];
      float* r = buf[cur ^ 1];
      nr = 0;
      for (int i = 0; i < nq; i++) {
        const float* pq = q + 2 * i;
        const float* nx = q + 2 * ((i + 1) % nq);
        bool in0 = sign * pq[dir] < h[dir], in1 = sign * nx[dir] < h[dir];
        if (in0) {
          r[2 * nr] = pq[0]; r[2 * nr + 1] = pq[1]; nr++;
          if (nr & 8) { full = true; break; }
        }
        if (in0 != in1) {
          r[2 * nr + (1 - dir)] = pq[1 - dir] + (nx[1 - dir] - pq[1 - dir]) / (nx[dir] - pq[dir]) * (sign * h[dir] - pq[dir]);
          r[2 * nr + dir] = sign * h[dir];
          nr++;
          if (nr & 8) { full = true; break; }
        }
      }
      cur ^= 1;
      nq = nr;
    }
  }
  for (int i = 0; i < 2 * nr; i++) out[i] = buf[cur][i];
  return nr;
}

__device__ void cull_points(int n, const float* p, int m, int i0, int* iret, float* A, bool* avail) {
  float a, cx, cy, q;
  if (n == 1) { cx = p[0]; cy = p[1]; }
  else if (n == 2) { cx = 0.5f * (p[0] + p[2]); cy = 0.5f * (p[1] + p[3]); }
  else {
    a = 0; cx = 0; cy = 0;
    for (int i = 0; i < n - 1; i++) {
      q = p[2 * i] * p[2 * i + 3] - p[2 * i + 2] * p[2 * i + 1];
      a += q; cx += q * (p[2 * i] + p[2 * i + 2]); cy += q * (p[2 * i + 1] + p[2 * i + 3]);
    }
    q = p[2 * n - 2] * p[1] - p[0] * p[2 * n - 1];
    a = fabsf(a + q) > FLT_EPSILON ? 1.0f / (3.0f * (a + q)) : 1e18f;
    cx = a * (cx + q * (p[2 * n - 2] + p[0]));
    cy = a * (cy + q * (p[2 * n - 1] + p[1]));
  }
  for (int i = 0; i < n; i++) { A[i] = atan2f(p[2 * i + 1] - cy, p[2 * i] - cx); avail[i] = true; }
  avail[i0] = false; iret[0] = i0;
  for (int j = 1; j < m; j++) {
    a = j * (2.0f * PI_F / m) + A[i0];
    if (a > PI_F) a -= 2.0f * PI_F;
    float best = 1e9f; int bi = i0;
    for (int i = 0; i < n; i++) if (avail[i]) {
      float d = fabsf(A[i] - a);
      if (d > PI_F) d = 2.0f * PI_F - d;
      if (d < best) { best = d; bi = i; }
    }
    avail[bi] = false; iret[j] = bi;
  }
}

// Boxes: centre p, orientation R (columns = box axes), half extents.  Up to 4 contacts out:
// point on B, normal on B (pointing from B to A), signed distance (<= 0).
//
// stat: 1 / 2 = box 1 / box 2 is a STATIC, AXIS-ALIGNED box (the table, the floor; R = identity) -- 0 = no promise.
// When the other box D lies over the static box S's top face, well inside its outline, the separating-axis search
// is decided before it starts: with m = the distance of D's centre from the nearest side of S's top face (in the
// plane), r = D's half diagonal and delta = the penetration along z, every unit axis n has
//   overlap(n) >= m (|nx| + |ny|) - (z_D - z_top) |nz| + support_D(n) >= (m - r) |n_xy| + delta |nz| >= delta
// as soon as m - r >= delta (support_D(n) >= |x* . n| for D's lowest point x*, |x*_xy| <= r, 1 - |nz| <= |n_xy|).
// So S's two side axes and the nine edge-edge axes (which must even beat the best face by the 1.05 fudge factor)
// cannot win; they are skipped.  The four remaining face axes (S's z, D's three) are evaluated with the general
// path's arithmetic, in its order, so the winner -- ties included -- and every contact are bit-identical to stat = 0
// (tests/test_coop_emu.py::test_box_box_static_fast_path_is_bit_identical).  This is the usual contact of the path:
// a finger or a block resting on the table.
__device__ int box_box(V3 p1, const M3& R1, V3 A, V3 p2, const M3& R2, V3 B, BoxScratch& scr, int stat = 0) {
  Contact* out = scr.out;
  V3 p = p2 - p1;
  V3 pp = mulT(R1, p);
  bool top = false;  // the fast path applies
  if (stat) {
    const V3 c = stat == 1 ? p : -p;                 // D's centre relative to S's (world axes = S's axes)
    const V3 S = stat == 1 ? A : B, D = stat == 1 ? B : A;
    const M3& RD = stat == 1 ? R2 : R1;
    const float rz = D.x * fabsf(RD.r2.x) + D.y * fabsf(RD.r2.y) + D.z * fabsf(RD.r2.z);   // D's extent along z
    const float delta = rz - (c.z - S.z);            // penetration along z (< 0: separated, the z axis test returns)
    const float m = fminf(S.x - fabsf(c.x), S.y - fabsf(c.y));
    top = c.z > 0.0f && m - norm(D) >= fmaxf(delta, 0.0f) + 1e-3f;
  }
  const bool skip1 = top && stat == 1, skip2 = top && stat == 2;
  float R[3][3], Q[3][3];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    V3 ci = col(R1, i);
#pragma unroll
    for (int j = 0; j < 3; j++) { R[i][j] = dot(ci, col(R2, j)); Q[i][j] = fabsf(R[i][j]); }
  }
  const float Ah[3] = {A.x, A.y, A.z}, Bh[3] = {B.x, B.y, B.z}, ppa[3] = {pp.x, pp.y, pp.z};
  float s = -FLT_MAX, s2;
  int code = 0; bool invert = false;
  int from_R = 0, ncol = 0;
  V3 normalC = v3(0, 0, 0);
#pragma unroll
  for (int i = 0; i < 3; i++) {
    if (i < 2 && skip1) continue;
    s2 = fabsf(ppa[i]) - (Ah[i] + Bh[0] * Q[i][0] + Bh[1] * Q[i][1] + Bh[2] * Q[i][2]);
    if (s2 > 0) return 0;
    if (s2 > s) { s = s2; from_R = 1; ncol = i; invert = ppa[i] < 0; code = i + 1; }
  }
#pragma unroll
  for (int j = 0; j < 3; j++) {
    if (j < 2 && skip2) continue;
    float e1 = dot(col(R2, j), p);
    s2 = fabsf(e1) - (Ah[0] * Q[0][j] + Ah[1] * Q[1][j] + Ah[2] * Q[2][j] + Bh[j]);
    if (s2 > 0) return 0;
    if (s2 > s) { s = s2; from_R = 2; ncol = j; invert = e1 < 0; code = j + 4; }
  }
  if (!top) {
#pragma unroll
  for (int i = 0; i < 3; i++) {
#pragma unroll
    for (int j = 0; j < 3; j++) Q[i][j] += 1.0e-5f;
  }
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
#pragma unroll
    for (int j = 0; j < 3; j++) {
      const int j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      float e1 = ppa[i2] * R[i1][j] - ppa[i1] * R[i2][j];
      float e2 = Ah[i1] * Q[i2][j] + Ah[i2] * Q[i1][j] + Bh[j1] * Q[i][j2] + Bh[j2] * Q[i][j1];
      s2 = fabsf(e1) - e2;
      if (s2 > FLT_EPSILON) return 0;
      float n1 = -R[i2][j], n2 = R[i1][j];
      float l = sqrtf(n1 * n1 + n2 * n2);
      if (l > FLT_EPSILON) {
        s2 /= l;
        if (s2 * BOX_FUDGE > s) {
          s = s2; from_R = 0;
          float nn[3]; nn[i] = 0; nn[i1] = n1 / l; nn[i2] = n2 / l;
          normalC = v3(nn[0], nn[1], nn[2]);
          invert = e1 < 0; code = 7 + 3 * i + j;
        }
      }
    }
  }
  }  // !top
  if (!code) return 0;
  V3 normal = from_R == 1 ? col(R1, ncol) : (from_R == 2 ? col(R2, ncol) : mul(R1, normalC));
  if (invert) normal = -normal;
  float depth = -s;
  if (code > 6) {  // edge-edge: the closest point on B's edge
    V3 pa = p1, pb = p2;
#pragma unroll
    for (int j = 0; j < 3; j++) {
      V3 c1 = col(R1, j), c2 = col(R2, j);
      pa += ((dot(normal, c1) > 0 ? 1.0f : -1.0f) * Ah[j]) * c1;
      pb += ((dot(normal, c2) > 0 ? -1.0f : 1.0f) * Bh[j]) * c2;
    }
    V3 ua = col(R1, (code - 7) / 3), ub = col(R2, (code - 7) % 3);
    V3 d = pb - pa;
    float uaub = dot(ua, ub), q1 = dot(ua, d), q2 = -dot(ub, d), den = 1 - uaub * uaub;
    float beta = den <= 1e-4f ? 0.0f : (uaub * q1 + q2) / den;
    out[0].pB = pb + beta * ub; out[0].nB = -normal; out[0].dist = -depth;
    return 1;
  }
  // face contact: reference face on `a`, incident face on `b`
  const bool swap = code > 3;
  M3 Ra, Rb;  // by value, with selects: a reference picked at run time would pin R1 / R2 to local memory
  Ra.r0 = swap ? R2.r0 : R1.r0; Ra.r1 = swap ? R2.r1 : R1.r1; Ra.r2 = swap ? R2.r2 : R1.r2;
  Rb.r0 = swap ? R1.r0 : R2.r0; Rb.r1 = swap ? R1.r1 : R2.r1; Rb.r2 = swap ? R1.r2 : R2.r2;
  V3 pa = swap ? p2 : p1, pb = swap ? p1 : p2;
  // half extents of the reference / incident box, picked with selects: indexing Ah / Bh dynamically would move
  // both arrays (and every SAT read of them above) to local memory
  const V3 SaV = swap ? B : A, SbV = swap ? A : B;
  V3 normal2 = swap ? -normal : normal;
  V3 nr = mulT(Rb, normal2);
  float anr[3] = {fabsf(nr.x), fabsf(nr.y), fabsf(nr.z)};
  int lanr, a1, a2;
  if (anr[1] > anr[0]) {
    if (anr[1] > anr[2]) { a1 = 0; lanr = 1; a2 = 2; } else { a1 = 0; a2 = 1; lanr = 2; }
  } else {
    if (anr[0] > anr[2]) { lanr = 0; a1 = 1; a2 = 2; } else { a1 = 0; a2 = 1; lanr = 2; }
  }
  V3 center = pb - pa + ((comp(nr, lanr) < 0 ? 1.0f : -1.0f) * comp(SbV, lanr)) * col(Rb, lanr);
  int codeN = swap ? code - 4 : code - 1, code1, code2;
  if (codeN == 0) { code1 = 1; code2 = 2; } else if (codeN == 1) { code1 = 0; code2 = 2; } else { code1 = 0; code2 = 1; }
  V3 ra1 = col(Ra, code1), ra2 = col(Ra, code2), rb1 = col(Rb, a1), rb2 = col(Rb, a2);
  float c1 = dot(center, ra1), c2 = dot(center, ra2);
  // Depths and contact points are evaluated relative to the centre of the REFERENCE FACE, not of the reference
  // box: the large terms (half extent against centre distance) cancel once, here, with a rounding error that is
  // common to all corners of the contact, and the per-corner terms are then added to a small number.  With
  // everything relative to the box centre the four corners of a block resting on the table differ by fp32
  // rounding of a 0.08 m coordinate (7e-9 m), which the contact ERP (450 1/s) turns into visible spin.
  const float sa_n = comp(SaV, codeN);
  const V3 centerf = center - sa_n * normal2;   // incident face centre relative to the reference face centre
  const V3 face_c = pa + sa_n * normal2;        // reference face centre, in the caller's coordinates
  float m11 = dot(ra1, rb1), m12 = dot(ra1, rb2), m21 = dot(ra2, rb1), m22 = dot(ra2, rb2);
  float* quad = scr.quad;
  float qx[4], qy[4];  // the incident face's corners in the reference face's 2-D frame (registers)
  {
    const float sb1 = comp(SbV, a1), sb2 = comp(SbV, a2);
    float k1 = m11 * sb1, k2 = m21 * sb1, k3 = m12 * sb2, k4 = m22 * sb2;
    qx[0] = c1 - k1 - k3; qy[0] = c2 - k2 - k4;
    qx[1] = c1 - k1 + k3; qy[1] = c2 - k2 + k4;
    qx[2] = c1 + k1 + k3; qy[2] = c2 + k2 + k4;
    qx[3] = c1 + k1 - k3; qy[3] = c2 + k2 - k4;
  }
  {
    // Fast path, the usual resting contact: the incident face lies inside the reference face (nothing to
    // clip) and all four corners touch (nothing to cull): the contacts are the four corners, in order.  The
    // arithmetic per corner is the generic path's, so the output is bit-identical to it.
    const float ra = comp(SaV, code1), rb = comp(SaV, code2);
    bool all_in = true;
#pragma unroll
    for (int j = 0; j < 4; j++) all_in = all_in && fabsf(qx[j]) < ra && fabsf(qy[j]) < rb;
    if (all_in) {
      const float det = 1.0f / (m11 * m22 - m12 * m21);
      const float i11 = m11 * det, i12 = m12 * det, i21 = m21 * det, i22 = m22 * det;
      V3 pt[4]; float dp[4];
      bool all_touch = true;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        float k1 = i22 * (qx[j] - c1) - i12 * (qy[j] - c2);
        float k2 = -i21 * (qx[j] - c1) + i11 * (qy[j] - c2);
        pt[j] = centerf + k1 * rb1 + k2 * rb2;
        dp[j] = -dot(normal2, pt[j]);
        all_touch = all_touch && dp[j] >= 0;
      }
      if (all_touch) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
          V3 w = pt[j] + face_c;
          if (swap) w -= dp[j] * normal;  // incident face was on A: move the point onto B
          out[j].pB = w; out[j].nB = -normal; out[j].dist = -dp[j];
        }
        return 4;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; j++) { quad[2 * j] = qx[j]; quad[2 * j + 1] = qy[j]; }
  float rect[2] = {comp(SaV, code1), comp(SaV, code2)};
  float* ret = scr.ret;
  int n = clip_quad_to_rect(rect, quad, ret, scr.buf);
  if (n < 1) return 0;
  V3* point = scr.point; float* dep = scr.dep;
  float det1 = 1.0f / (m11 * m22 - m12 * m21);
  m11 *= det1; m12 *= det1; m21 *= det1; m22 *= det1;
  int cnum = 0;
  for (int j = 0; j < n; j++) {
    float k1 = m22 * (ret[2 * j] - c1) - m12 * (ret[2 * j + 1] - c2);
    float k2 = -m21 * (ret[2 * j] - c1) + m11 * (ret[2 * j + 1] - c2);
    V3 pt = centerf + k1 * rb1 + k2 * rb2;
    float dp = -dot(normal2, pt);
    if (dp >= 0) { point[cnum] = pt; dep[cnum] = dp; ret[2 * cnum] = ret[2 * j]; ret[2 * cnum + 1] = ret[2 * j + 1]; cnum++; }
  }
  if (cnum < 1) return 0;
  int maxc = cnum < 4 ? cnum : 4;
  int* idx = scr.idx;
  if (cnum <= maxc) { for (int j = 0; j < cnum; j++) idx[j] = j; }
  else {
    int i1 = 0; float maxdepth = dep[0];
    for (int i = 1; i < cnum; i++) if (dep[i] > maxdepth) { maxdepth = dep[i]; i1 = i; }
    cull_points(cnum, ret, maxc, i1, idx, scr.ang, scr.avail);
  }
  for (int j = 0; j < maxc; j++) {
    int k = idx[j];
    V3 w = point[k] + face_c;
    if (swap) w -= dep[k] * normal;  // incident face was on A: move the point onto B
    out[j].pB = w; out[j].nB = -normal; out[j].dist = -dep[k];
  }
  return maxc;
}

// ---- box (A) against a cylinder (B, axis = its local z): the Slide puck -----------------------------------------
// A model of the pair, not a restatement of Bullet's GJK/EPA path (see DESIGN.md on the puck model,
// Slide): separating-axis choice between the cylinder axis (a cap against a box face) and the radial direction (the
// curved side against the box), then
//   cap - face : the rim points of that cap at four azimuths fixed in the cylinder (+-x, +-y), plus the deepest rim
//                point once the cap is tilted by more than ~1 degree, kept when they lie over the face and penetrate
//                it, and the corners of the face inside the cap disc; normal = the face normal; <= 4 points (the
//                deepest, then the ones farthest from those already kept);
//   side       : the closest points of the box to the cylinder axis at the two ends of their common height range.
// Depths are taken relative to the centre of the box face (one rounding common to all points, as in box_box).
// Output convention of box_box: point on B, normal on B (pointing from B towards A), signed distance (<= 0).
__device__ int box_cyl(V3 pa, const M3& Ra, V3 ha, V3 pb, const M3& Rb, float r, float h, Contact* out) {
  const V3 cb = mulT(Rb, pa - pb);                           // box centre in the cylinder frame
  V3 Aq[3];                                                  // box axes in the cylinder frame
  Aq[0] = mulT(Rb, col(Ra, 0)); Aq[1] = mulT(Rb, col(Ra, 1)); Aq[2] = mulT(Rb, col(Ra, 2));
  const float hav[3] = {ha.x, ha.y, ha.z};
  const float ez = ha.x * fabsf(Aq[0].z) + ha.y * fabsf(Aq[1].z) + ha.z * fabsf(Aq[2].z);   // box extent along the axis
  const float scap = cb.z >= 0.0f ? 1.0f : -1.0f;            // the cap on the box's side
  const float sep_cap = scap * cb.z - ez - h;
  if (sep_cap > 0.0f) return 0;
  const float z0 = fmaxf(cb.z - ez, -h), z1 = fminf(cb.z + ez, h);
  const float zs[2] = {z0 + 0.05f * (z1 - z0), z1 - 0.05f * (z1 - z0)};
  V3 qs[2];
  float rho[2], sep_rad = 1e30f;
#pragma unroll
  for (int i = 0; i < 2; i++) {
    const V3 rel = v3(-cb.x, -cb.y, zs[i] - cb.z);
    V3 q = cb;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const float loc = fminf(fmaxf(dot(rel, Aq[k]), -hav[k]), hav[k]);   // clamp in the box frame
      q += loc * Aq[k];
    }
    qs[i] = q;
    rho[i] = sqrtf(q.x * q.x + q.y * q.y);
    sep_rad = fminf(sep_rad, rho[i] - r);
  }
  if (sep_rad > 0.0f) return 0;
  int n = 0;
  if (sep_cap >= sep_rad) {
    // ---- cap against the box face whose outward normal opposes the cap normal most ----
    int fj = 0;
    float best = fabsf(Aq[0].z);
    if (fabsf(Aq[1].z) > best) { best = fabsf(Aq[1].z); fj = 1; }
    if (fabsf(Aq[2].z) > best) { best = fabsf(Aq[2].z); fj = 2; }
    const int k1 = (fj + 1) % 3, k2 = (fj + 2) % 3;
    const V3 af = fj == 0 ? Aq[0] : (fj == 1 ? Aq[1] : Aq[2]);
    const V3 a1 = k1 == 0 ? Aq[0] : (k1 == 1 ? Aq[1] : Aq[2]), a2 = k2 == 0 ? Aq[0] : (k2 == 1 ? Aq[1] : Aq[2]);
    const float hf = hav[fj], h1 = hav[k1], h2 = hav[k2];
    const float sig = af.z * scap > 0.0f ? -1.0f : 1.0f;     // the face looks against the cap normal
    const V3 nf = sig * af;                                  // outward face normal, cylinder frame
    const V3 fc = cb + (sig * hf) * af;                      // centre of that face
    float cand[9][4];                                        // point on B (cylinder frame), distance
    int nc = 0;
    const float tx = -nf.x, ty = -nf.y, tn = sqrtf(tx * tx + ty * ty);
    const int nu = tn > 0.02f ? 5 : 4;                       // tilted by more than ~1 degree: the lowest rim point matters
#pragma unroll 1
    for (int i = 0; i < nu; i++) {
      const float ux = i == 0 ? 1.0f : (i == 1 ? -1.0f : (i < 4 ? 0.0f : tx / tn));
      const float uy = i < 2 ? 0.0f : (i == 2 ? 1.0f : (i == 3 ? -1.0f : ty / tn));
      const V3 p = v3(r * ux, r * uy, scap * h), rel = p - fc;
      if (fabsf(dot(rel, a1)) > h1 || fabsf(dot(rel, a2)) > h2) continue;
      const float dist = dot(rel, nf);
      if (dist > 0.0f) continue;
      cand[nc][0] = p.x; cand[nc][1] = p.y; cand[nc][2] = p.z; cand[nc][3] = dist;
      nc++;
    }
    const float ncn = -(nf.z * scap);                        // n . cap normal with n = -nf
    if (ncn > 1e-6f) {
#pragma unroll 1
      for (int i = 0; i < 4; i++) {                          // corners of the face inside the cap disc
        const V3 v = fc + ((i & 1) ? h1 : -h1) * a1 + ((i & 2) ? h2 : -h2) * a2;
        const float dist = (v.z - scap * h) * scap / ncn;    // along n = -nf
        if (dist > 0.0f) continue;
        const V3 pB = v + dist * nf;                         // pB = pA - dist n
        if (pB.x * pB.x + pB.y * pB.y > r * r) continue;
        cand[nc][0] = pB.x; cand[nc][1] = pB.y; cand[nc][2] = pB.z; cand[nc][3] = dist;
        nc++;
      }
    }
    int keep[4];
    unsigned used = 0;
#pragma unroll 1
    for (int m = 0; m < 4 && m < nc; m++) {
      int bi = -1;
      float bv = -1e30f;
#pragma unroll 1
      for (int i = 0; i < nc; i++) {
        if (used & (1u << i)) continue;
        float score;
        if (m == 0) score = -cand[i][3];
        else {
          score = 1e30f;
          for (int j = 0; j < m; j++) {
            const float e0 = cand[i][0] - cand[keep[j]][0], e1 = cand[i][1] - cand[keep[j]][1], e2 = cand[i][2] - cand[keep[j]][2];
            score = fminf(score, e0 * e0 + e1 * e1 + e2 * e2);
          }
          if (score < 1e-10f) continue;                      // coincides with a kept point
        }
        if (score > bv) { bv = score; bi = i; }
      }
      if (bi < 0) break;
      used |= 1u << bi; keep[m] = bi;
      out[n].pB = pb + mul(Rb, v3(cand[bi][0], cand[bi][1], cand[bi][2]));
      out[n].nB = mul(Rb, -nf);
      out[n].dist = cand[bi][3];
      n++;
    }
    return n;
  }
  // ---- curved side against the box ----
#pragma unroll
  for (int i = 0; i < 2; i++) {
    if (rho[i] - r > 0.0f || rho[i] < 1e-9f) continue;
    if (i == 1 && fabsf(zs[1] - zs[0]) < 1e-6f) break;
    const V3 nl = v3(qs[i].x / rho[i], qs[i].y / rho[i], 0.0f);
    out[n].pB = pb + mul(Rb, v3(nl.x * r, nl.y * r, qs[i].z));
    out[n].nB = mul(Rb, nl);
    out[n].dist = rho[i] - r;
    n++;
  }
  return n;
}

// Cylinder A (axis = its local z) against box B: box_cyl with the roles exchanged, its contacts turned round to this
// file's convention (point on B, normal on B pointing from B towards A).  The gripper base against a block.
// Out of line: it runs only while the base is within reach of a block, and its code stays out of the narrowphase loop.
__device__ __noinline__ int cyl_box(V3 pa, const M3& Ra, float r, float h, V3 pb, const M3& Rb, V3 hb, Contact* out) {
  const int n = box_cyl(pb, Rb, hb, pa, Ra, r, h, out);
  for (int i = 0; i < n; i++) {
    out[i].pB = out[i].pB + out[i].dist * out[i].nB;  // point on the box = point on the cylinder + n * distance
    out[i].nB = -out[i].nB;
  }
  return n;
}

__device__ __forceinline__ void plane_space(V3 n, V3& p, V3& q) {  // btPlaneSpace1
  if (fabsf(n.z) > 0.70710678118654752440f) {
    float a = n.y * n.y + n.z * n.z, k = 1.0f / sqrtf(a);
    p = v3(0, -n.z * k, n.y * k);
    q = v3(a * k, -n.x * p.z, n.x * p.y);
  } else {
    float a = n.x * n.x + n.y * n.y, k = 1.0f / sqrtf(a);
    p = v3(-n.y * k, n.x * k, 0);
    q = v3(-n.z * p.y, n.z * p.x, a * k);
  }
}

}  // namespace pmg
