/* pmg_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  See pmg_oracle.h for the contract.
 *
 * PARITY UNPINNED (physics): restated from the published Bullet algorithms as recollected,
 * no pybullet available to confirm.  Reference call sites this file follows
 * (paths relative to /root/reference/pybullet_multigoal_gym/):
 *   envs/base_envs/base_env.py:124-138,215-219      reset / step / physics parameters
 *   robots/kuka.py:27-51,120-172,207-301            rest pose, bounds, action map, IK call, motors
 *   robots/robot_bases.py:108-133,230-238           getLinkState semantics, reset_position
 *   envs/base_envs/kuka_single_step_base_env.py:104-148,193-244   sampling, obs, reward
 *   envs/base_envs/kuka_multi_step_base_env.py:223-246,255-345    multi-block sampling, obs, reward
 *   envs/task_envs/kuka_multi_step_envs.py:34-87    block-stack goal
 * Third-party arithmetic restated (pybullet ~= 3.0.6, requirements.txt:2):
 *   btMultiBody ABA + unit-impulse responses, btMultiBodyConstraintSolver (PGS, 5 iterations),
 *   btMultiBodyJointMotor / JointLimitConstraint rows, btBoxBoxDetector, btPersistentManifold,
 *   IKTrajectoryHelper + BussIK damped least squares, numpy legacy RandomState (MT19937).
 */
#include "pmg_oracle.h"

#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "../include/pmg_model_constants.h"

/* ------------------------------------------------------------------------------------------ */
/* physics parameters (base_env.py:215-219, kuka.py:223-225,282-301; Bullet defaults B.1)     */
/* ------------------------------------------------------------------------------------------ */
#define DT 0.002            /* fixedTimeStep 0.04 / numSubSteps 20 */
#define OUTER_DT 0.04       /* m_physicsDeltaTime: scales the motor max impulse [BULLET-MEMORY] */
#define SUBSTEPS_PER_CALL 20
#define CALLS_PER_ENV_STEP 5 /* kuka.py:223-225 */
#define SOLVER_ITERS 5       /* base_env.py:37,218 */
#define CONTACT_ERP 0.9      /* setDefaultContactERP, base_env.py:216 -> m_erp2 */
#define LINEAR_SLOP 1e-5
#define RESIDUAL_THRESHOLD 1e-7 /* m_leastSquaresResidualThreshold set by pybullet */
#define GRAVITY_Z (-9.81)
#define LINK_DAMPING 0.04     /* btMultiBody m_linearDamping = m_angularDamping */
#define MAX_COORD_VEL 100.0   /* btMultiBody m_maxCoordinateVelocity */
#define LIMIT_MAX_IMPULSE 100.0 /* btMultiBodyConstraint default m_maxAppliedImpulse */
#define SPLIT_IMPULSE_PEN_THRESHOLD (-0.04)
#define MOTOR_KP 0.03
#define MOTOR_KD 1.0
#define ARM_FORCE 200.0
#define FINGER_FORCE 50.0
#define BREAKING_THRESHOLD_FACTOR 0.02 /* gContactBreakingThreshold, relative to the shape radius */
#define BROADPHASE_MARGIN 0.02
#define IK_JOINT_DAMPING 0.5
#define IK_MAX_STEP (45.0 * M_PI / 180.0)
#define BOX_FUDGE 1.05

#define NB PMG_NBODY
#define ND PMG_NDOF
#define MAXBLK PMGO_MAX_BLOCKS
#define MAX_PAIRS (2 + 4 * MAXBLK + MAXBLK * (MAXBLK - 1) / 2 + MAXBLK)
#define MAX_ROWS (2 * ND + ND + 3 * 4 * MAX_PAIRS)

static const int B_PARENT[NB] = PMG_BODY_PARENT;
static const int B_JTYPE[NB] = PMG_BODY_JTYPE;
static const int B_DOF[NB] = PMG_BODY_DOF;
static const double B_JXYZ[NB][3] = PMG_BODY_JXYZ;
static const double B_JROT[NB][9] = PMG_BODY_JROT;
static const double B_AXIS[NB][3] = PMG_BODY_AXIS;
static const double B_MASS[NB] = PMG_BODY_MASS;
static const double B_COM[NB][3] = PMG_BODY_COM;
static const double B_INERTIA[NB][3] = PMG_BODY_INERTIA;
static const double DOF_LOWER[ND] = PMG_DOF_LOWER;
static const double DOF_UPPER[ND] = PMG_DOF_UPPER;
static const double DOF_DAMPING[ND] = PMG_DOF_DAMPING;
static const int DOF_BODY[ND] = PMG_DOF_BODY;
static const double TIP_OFFSET[3] = PMG_TIP_OFFSET;
static const double TAB1_OFFSET[3] = PMG_TAB1_OFFSET;
static const double TAB2_OFFSET[3] = PMG_TAB2_OFFSET;
static const double FINGER_HALF[3] = PMG_FINGER_HALF;
static const double TABLE_CENTER[3] = PMG_TABLE_CENTER;
static const double TABLE_HALF[3] = PMG_TABLE_HALF;
static const double FLOOR_CENTER[3] = PMG_FLOOR_CENTER;
static const double FLOOR_HALF[3] = PMG_FLOOR_HALF;
static const int NONCONTACT_ORDER[2 * ND] = PMG_NONCONTACT_ORDER;
/* Slide (kuka_single_step_envs.py:49-59, kuka_single_step_base_env.py:53-56,66-69,89-93) */
static const double LONG_TABLE_CENTER[3] = PMG_LONG_TABLE_CENTER;
static const double LONG_TABLE_HALF[3] = PMG_LONG_TABLE_HALF;
static const double PUCK_INERTIA[3] = PMG_PUCK_INERTIA;

/* kuka.py:27 */
static const double KUKA_REST_POSE[7] = {0, -0.5592432, 0, 1.733180, 0, -0.8501557, 0};
/* kuka.py:40-42 */
static const double EE_UPPER[3] = {-0.37, 0.20, 0.55};
static const double EE_LOWER[3] = {-0.67, -0.20, 0.175};
static const double EE_FIXED_QUAT[4] = {0, -1, 0, 0};
#define GRIPPER_ABS_LIMIT 0.035 /* kuka.py:71 */
#define BLOCK_SPAWN_Z 0.175     /* kuka_single_step_base_env.py:50 */

/* ------------------------------------------------------------------------------------------ */
/* small vector helpers                                                                       */
/* ------------------------------------------------------------------------------------------ */
static inline double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline void cross3(const double* a, const double* b, double* o) {
  double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  o[0] = x; o[1] = y; o[2] = z;
}
static inline void set3(double* o, double x, double y, double z) { o[0] = x; o[1] = y; o[2] = z; }
static inline void copy3(double* o, const double* a) { o[0] = a[0]; o[1] = a[1]; o[2] = a[2]; }
static inline void add3(double* o, const double* a, const double* b) { o[0] = a[0] + b[0]; o[1] = a[1] + b[1]; o[2] = a[2] + b[2]; }
static inline void sub3(double* o, const double* a, const double* b) { o[0] = a[0] - b[0]; o[1] = a[1] - b[1]; o[2] = a[2] - b[2]; }
static inline void axpy3(double* o, double s, const double* a) { o[0] += s * a[0]; o[1] += s * a[1]; o[2] += s * a[2]; }
static inline double norm3(const double* a) { return sqrt(dot3(a, a)); }
/* row-major 3x3 */
static inline void matvec3(const double* R, const double* v, double* o) {
  double x = R[0] * v[0] + R[1] * v[1] + R[2] * v[2];
  double y = R[3] * v[0] + R[4] * v[1] + R[5] * v[2];
  double z = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}
static inline void matTvec3(const double* R, const double* v, double* o) {
  double x = R[0] * v[0] + R[3] * v[1] + R[6] * v[2];
  double y = R[1] * v[0] + R[4] * v[1] + R[7] * v[2];
  double z = R[2] * v[0] + R[5] * v[1] + R[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}
static inline void matmul3(const double* A, const double* B, double* C) {
  double t[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) t[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
  memcpy(C, t, sizeof t);
}
static void axis_angle_mat(const double* a, double ang, double* R) {
  double c = cos(ang), s = sin(ang), t = 1 - c, x = a[0], y = a[1], z = a[2];
  R[0] = t * x * x + c;     R[1] = t * x * y - s * z; R[2] = t * x * z + s * y;
  R[3] = t * x * y + s * z; R[4] = t * y * y + c;     R[5] = t * y * z - s * x;
  R[6] = t * x * z - s * y; R[7] = t * y * z + s * x; R[8] = t * z * z + c;
}
static void quat_to_mat(const double* q, double* R) { /* xyzw */
  double x = q[0], y = q[1], z = q[2], w = q[3];
  double d = x * x + y * y + z * z + w * w, s = 2.0 / d;
  double xs = x * s, ys = y * s, zs = z * s;
  double wx = w * xs, wy = w * ys, wz = w * zs, xx = x * xs, xy = x * ys, xz = x * zs, yy = y * ys, yz = y * zs, zz = z * zs;
  R[0] = 1 - (yy + zz); R[1] = xy - wz; R[2] = xz + wy;
  R[3] = xy + wz; R[4] = 1 - (xx + zz); R[5] = yz - wx;
  R[6] = xz - wy; R[7] = yz + wx; R[8] = 1 - (xx + yy);
}
static void mat_to_quat(const double* m, double* q) { /* btMatrix3x3::getRotation, xyzw */
  double trace = m[0] + m[4] + m[8];
  if (trace > 0) {
    double s = sqrt(trace + 1.0);
    q[3] = s * 0.5; s = 0.5 / s;
    q[0] = (m[7] - m[5]) * s; q[1] = (m[2] - m[6]) * s; q[2] = (m[3] - m[1]) * s;
  } else {
    int i = m[0] < m[4] ? (m[4] < m[8] ? 2 : 1) : (m[0] < m[8] ? 2 : 0);
    int j = (i + 1) % 3, k = (i + 2) % 3;
    double s = sqrt(m[4 * i] - m[4 * j] - m[4 * k] + 1.0);
    q[i] = s * 0.5; s = 0.5 / s;
    q[3] = (m[3 * k + j] - m[3 * j + k]) * s;
    q[j] = (m[3 * j + i] + m[3 * i + j]) * s;
    q[k] = (m[3 * k + i] + m[3 * i + k]) * s;
  }
}
static void quat_mul(const double* a, const double* b, double* o) { /* xyzw */
  double x = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
  double y = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
  double z = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
  double w = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
  o[0] = x; o[1] = y; o[2] = z; o[3] = w;
}

/* spatial vectors: [angular(3); linear(3)], all expressed in world axes at the world origin */
static inline double dot6(const double* a, const double* b) { return dot3(a, b) + dot3(a + 3, b + 3); }
static void crm(const double* v, const double* m, double* o) { /* motion cross: v x m */
  double a[3], b[3], c[3];
  cross3(v, m, a); cross3(v, m + 3, b); cross3(v + 3, m, c);
  copy3(o, a); add3(o + 3, b, c);
}
static void crf(const double* v, const double* f, double* o) { /* force cross: v x* f */
  double a[3], b[3], c[3];
  cross3(v, f, a); cross3(v + 3, f + 3, b); cross3(v, f + 3, c);
  add3(o, a, b); copy3(o + 3, c);
}
static void mat6vec(const double* M, const double* v, double* o) {
  double t[6];
  for (int i = 0; i < 6; i++) { double s = 0; for (int j = 0; j < 6; j++) s += M[6 * i + j] * v[j]; t[i] = s; }
  memcpy(o, t, sizeof t);
}

/* ------------------------------------------------------------------------------------------ */
/* numpy legacy RandomState (MT19937) restatement                                             */
/* ------------------------------------------------------------------------------------------ */
typedef struct { uint32_t mt[624]; int idx; } MT;
static void mt_init_genrand(MT* s, uint32_t seed) {
  s->mt[0] = seed;
  for (int i = 1; i < 624; i++) s->mt[i] = 1812433253u * (s->mt[i - 1] ^ (s->mt[i - 1] >> 30)) + (uint32_t)i;
  s->idx = 624;
}
static void mt_init_by_array(MT* s, const uint32_t* key, int len) {
  mt_init_genrand(s, 19650218u);
  int i = 1, j = 0, k = 624 > len ? 624 : len;
  for (; k; k--) {
    s->mt[i] = (s->mt[i] ^ ((s->mt[i - 1] ^ (s->mt[i - 1] >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
    i++; j++;
    if (i >= 624) { s->mt[0] = s->mt[623]; i = 1; }
    if (j >= len) j = 0;
  }
  for (k = 623; k; k--) {
    s->mt[i] = (s->mt[i] ^ ((s->mt[i - 1] ^ (s->mt[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
    i++;
    if (i >= 624) { s->mt[0] = s->mt[623]; i = 1; }
  }
  s->mt[0] = 0x80000000u;
  s->idx = 624;
}
static uint32_t mt_next(MT* s) {
  if (s->idx >= 624) {
    uint32_t* mt = s->mt;
    for (int k = 0; k < 624; k++) {
      uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
      mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    s->idx = 0;
  }
  uint32_t y = s->mt[s->idx++];
  y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
  return y;
}
static double mt_double(MT* s) { /* random_sample: 53-bit */
  uint32_t a = mt_next(s) >> 5, b = mt_next(s) >> 6;
  return (a * 67108864.0 + b) / 9007199254740992.0;
}
static double mt_uniform(MT* s, double lo, double hi) { return lo + (hi - lo) * mt_double(s); }
static uint32_t mt_interval(MT* s, uint32_t max) { /* legacy random_interval (masked rejection) */
  if (max == 0) return 0;
  uint32_t mask = max, v;
  mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
  while ((v = (mt_next(s) & mask)) > max) {}
  return v;
}

/* ------------------------------------------------------------------------------------------ */
/* data structures                                                                            */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  int n;
  double lA[4][3], lB[4][3]; /* contact point in the local frames of A and B */
  double nB[4][3];           /* world normal on B (pointing from B towards A), as cached */
  double dist[4];
} Manifold;

/* collision endpoints: kind 0 static box, 1 robot body, 2 block */
typedef struct { int kind, index; double half[3]; double friction; double radius; int cylinder; } Geom; /* cylinder: axis z, half = (r, r, h) */
typedef struct { Geom a, b; } Pair;

typedef struct {
  /* robot part (body A when the robot takes part) */
  int has_robot; double Jr[ND], MJr[ND];
  int blkA, blkB;            /* block indices or -1 */
  double JA[6], MJA[6], JB[6], MJB[6]; /* [angular; linear] */
  double rhs, lo, hi, diag_inv, applied, mu;
  int normal_row;            /* friction rows: index of their normal row */
} Row;

struct PmgoEnv {
  int task, nb, binary, max_steps, grasping, has_obj, adim;
  int multi, grip, jc; /* multi-block obs layout (stack / rearrange); grip-informed goal; joint-space control */
  int td, sub_goal_ind; /* task decomposition (block_stack): which sub-goal is the desired goal, -1 = the last one */
  /* curriculum (block_stack, kuka_multi_step_base_env.py:122-140,350-379): goal difficulty drawn per reset */
  int cur, cur_update, cur_level; double cur_prob[MAXBLK], cur_count[MAXBLK], cur_goals_per;
  double thr;
  int dims[4];
  /* state */
  double q[ND], qd[ND];
  double bpos[MAXBLK][3], bquat[MAXBLK][4], bv[MAXBLK][3], bw[MAXBLK][3];
  double ee_target[3], rest_pose[7];
  double mot_target[ND], mot_maximp[ND];
  double joint_damp_tau[ND];
  double goal[3 * MAXBLK];
  int elapsed;
  /* multi-block goal bookkeeping (kuka_multi_step_envs.py:62-63) */
  int last_order[MAXBLK]; double last_targets[MAXBLK][3];
  double tip_init[3], obj_lo[3], obj_hi[3], tgt_lo[3], tgt_hi[3];
  int target_in_air;
  MT rng;
  /* contacts */
  int npair; Pair pairs[MAX_PAIRS]; Manifold man[MAX_PAIRS];
  /* free-body parameters (cubes: isotropic; the Slide puck: a cylinder about its z axis) and the scene's table */
  double blk_mass, blk_inertia[3], blk_spawn_z; int puck;
  double table_center[3];
  /* per-substep kinematics / ABA cache */
  double bR[NB][9], bp[NB][3], S[NB][6], U[NB][6], Dinv[NB], vel[NB][6];
  double blkR[MAXBLK][9];
  Row rows[MAX_ROWS];
};

/* ------------------------------------------------------------------------------------------ */
/* forward kinematics                                                                         */
/* ------------------------------------------------------------------------------------------ */
static void fk_bodies(const double* q, double bR[NB][9], double bp[NB][3], double S[NB][6]) {
  static const double I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  for (int b = 0; b < NB; b++) {
    const double* Rp = B_PARENT[b] < 0 ? I3 : bR[B_PARENT[b]];
    double pp[3] = {0, 0, 0};
    if (B_PARENT[b] >= 0) copy3(pp, bp[B_PARENT[b]]);
    double Rj[9], t[3], aw[3];
    matmul3(Rp, B_JROT[b], Rj);
    matvec3(Rp, B_JXYZ[b], t);
    add3(bp[b], pp, t);
    matvec3(Rj, B_AXIS[b], aw);
    if (B_JTYPE[b] == 0) {
      double Rq[9];
      axis_angle_mat(B_AXIS[b], q[B_DOF[b]], Rq);
      matmul3(Rj, Rq, bR[b]);
      if (S) { copy3(S[b], aw); cross3(bp[b], aw, S[b] + 3); }
    } else if (B_JTYPE[b] == 1) {
      memcpy(bR[b], Rj, sizeof(double) * 9);
      axpy3(bp[b], q[B_DOF[b]], aw);
      if (S) { set3(S[b], 0, 0, 0); copy3(S[b] + 3, aw); }
    } else {
      memcpy(bR[b], Rj, sizeof(double) * 9);
      if (S) memset(S[b], 0, sizeof(double) * 6);
    }
  }
}

void pmgo_fk_tip(const double q[9], double pos[3], double quat[4]) {
  double bR[NB][9], bp[NB][3], t[3];
  fk_bodies(q, bR, bp, NULL);
  matvec3(bR[PMG_BODY_LINK7], TIP_OFFSET, t);
  add3(pos, bp[PMG_BODY_LINK7], t);
  mat_to_quat(bR[PMG_BODY_LINK7], quat);
}

/* ------------------------------------------------------------------------------------------ */
/* inverse kinematics: pybullet calculateInverseKinematics without null space                 */
/* (kuka.py:266-279 passes 7-element null-space lists to a 9-DoF body => ignored, B.3)        */
/* ------------------------------------------------------------------------------------------ */
static void solve_dense(double* A, double* b, int n) { /* Gaussian elimination, partial pivoting */
  for (int c = 0; c < n; c++) {
    int piv = c;
    for (int r = c + 1; r < n; r++) if (fabs(A[r * n + c]) > fabs(A[piv * n + c])) piv = r;
    if (piv != c) {
      for (int k = 0; k < n; k++) { double t = A[c * n + k]; A[c * n + k] = A[piv * n + k]; A[piv * n + k] = t; }
      double t = b[c]; b[c] = b[piv]; b[piv] = t;
    }
    for (int r = c + 1; r < n; r++) {
      double f = A[r * n + c] / A[c * n + c];
      for (int k = c; k < n; k++) A[r * n + k] -= f * A[c * n + k];
      b[r] -= f * b[c];
    }
  }
  for (int r = n - 1; r >= 0; r--) {
    double s = b[r];
    for (int k = r + 1; k < n; k++) s -= A[r * n + k] * b[k];
    b[r] = s / A[r * n + r];
  }
}

void pmgo_ik(const double q_seed[9], const double target_pos[3], const double target_quat[4],
             int max_iter, double residual_threshold, double q_out[9]) {
  double q[ND];
  memcpy(q, q_seed, sizeof q);
  double diff = 1e30;
  for (int it = 0; it < max_iter && diff > residual_threshold; it++) {
    double bR[NB][9], bp[NB][3], S[NB][6], tip[3], t[3];
    fk_bodies(q, bR, bp, S);
    matvec3(bR[PMG_BODY_LINK7], TIP_OFFSET, t);
    add3(tip, bp[PMG_BODY_LINK7], t);
    /* 6 x 9 Jacobian of the tip frame; finger columns are zero */
    double J[6][ND];
    memset(J, 0, sizeof J);
    for (int b = 0; b <= PMG_BODY_LINK7; b++) {
      int d = B_DOF[b];
      double lin[3];
      /* v_tip = S_lin + S_ang x tip */
      cross3(S[b], tip, lin);
      add3(lin, lin, S[b] + 3);
      for (int k = 0; k < 3; k++) { J[k][d] = lin[k]; J[3 + k][d] = S[b][k]; }
    }
    double e[6];
    sub3(e, target_pos, tip);
    diff = norm3(e);
    /* orientation error: axis * angle of target * current^-1, angle stored as float (sic) */
    double qc[4], qinv[4], dq[4];
    mat_to_quat(bR[PMG_BODY_LINK7], qc);
    qinv[0] = -qc[0]; qinv[1] = -qc[1]; qinv[2] = -qc[2]; qinv[3] = qc[3];
    quat_mul(target_quat, qinv, dq);
    double w = dq[3] < -1 ? -1 : (dq[3] > 1 ? 1 : dq[3]);
    float angle = (float)(2.0 * acos(w));
    double axis[3], s2 = 1.0 - dq[3] * dq[3];
    if (s2 < 10.0 * DBL_EPSILON) set3(axis, 1, 0, 0);
    else { double s = 1.0 / sqrt(s2); set3(axis, dq[0] * s, dq[1] * s, dq[2] * s); }
    if (angle > (float)M_PI) angle -= (float)(2.0 * M_PI);
    else if (angle < -(float)M_PI) angle += (float)(2.0 * M_PI);
    double an = norm3(axis);
    for (int k = 0; k < 3; k++) e[3 + k] = (double)angle * axis[k] / an;
    /* (J^T J + diag(damping)) dtheta = J^T e */
    double A[ND * ND], rhs[ND];
    for (int i = 0; i < ND; i++) {
      for (int j = 0; j < ND; j++) {
        double s = 0;
        for (int k = 0; k < 6; k++) s += J[k][i] * J[k][j];
        A[i * ND + j] = s + (i == j ? IK_JOINT_DAMPING : 0.0);
      }
      double s = 0;
      for (int k = 0; k < 6; k++) s += J[k][i] * e[k];
      rhs[i] = s;
    }
    solve_dense(A, rhs, ND);
    double mx = 0;
    for (int i = 0; i < ND; i++) if (fabs(rhs[i]) > mx) mx = fabs(rhs[i]);
    if (mx > IK_MAX_STEP) for (int i = 0; i < ND; i++) rhs[i] *= IK_MAX_STEP / mx;
    for (int i = 0; i < ND; i++) q[i] += rhs[i];
  }
  memcpy(q_out, q, sizeof q);
}

/* ------------------------------------------------------------------------------------------ */
/* articulated-body algorithm (world-frame spatial algebra, reference point = world origin)   */
/* ------------------------------------------------------------------------------------------ */
static void body_world_inertia(const PmgoEnv* e, int b, double* com_w, double* Iw) {
  double t[3], RD[9];
  matvec3(e->bR[b], B_COM[b], t);
  add3(com_w, e->bp[b], t);
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) RD[3 * i + j] = e->bR[b][3 * i + j] * B_INERTIA[b][j];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      Iw[3 * i + j] = RD[3 * i] * e->bR[b][3 * j] + RD[3 * i + 1] * e->bR[b][3 * j + 1] + RD[3 * i + 2] * e->bR[b][3 * j + 2];
}

static void spatial_inertia_at_origin(double m, const double* c, const double* Ic, double* I6) {
  /* [[Ic - m c^ c^, m c^], [-m c^, m 1]] */
  double cx[9] = {0, -c[2], c[1], c[2], 0, -c[0], -c[1], c[0], 0};
  double cc[9];
  matmul3(cx, cx, cc);
  memset(I6, 0, sizeof(double) * 36);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      I6[6 * i + j] = Ic[3 * i + j] - m * cc[3 * i + j];
      I6[6 * i + 3 + j] = m * cx[3 * i + j];
      I6[6 * (3 + i) + j] = -m * cx[3 * i + j];
    }
  I6[21] = I6[28] = I6[35] = m;
}

/* Forward dynamics of the robot: q, qd, tau -> qdd.  Leaves S, U, Dinv, vel cached in e. */
static void aba(PmgoEnv* e, const double* tau, double* qdd) {
  double IA[NB][36], pA[NB][6], c[NB][6], u[NB], a[NB][6];
  fk_bodies(e->q, e->bR, e->bp, e->S);
  for (int b = 0; b < NB; b++) {
    double vJ[6] = {0, 0, 0, 0, 0, 0};
    if (B_DOF[b] >= 0) for (int k = 0; k < 6; k++) vJ[k] = e->S[b][k] * e->qd[B_DOF[b]];
    for (int k = 0; k < 6; k++) e->vel[b][k] = (B_PARENT[b] < 0 ? 0.0 : e->vel[B_PARENT[b]][k]) + vJ[k];
    crm(e->vel[b], vJ, c[b]);
    double com[3], Iw[9];
    body_world_inertia(e, b, com, Iw);
    spatial_inertia_at_origin(B_MASS[b], com, Iw, IA[b]);
    /* external force: gravity + Bullet's per-link velocity damping, applied at the COM */
    double w[3], vcom[3], F[3], T[3], Iww[3];
    copy3(w, e->vel[b]);
    cross3(w, com, vcom);
    add3(vcom, vcom, e->vel[b] + 3);
    double kl = LINK_DAMPING + LINK_DAMPING * norm3(vcom), ka = LINK_DAMPING + LINK_DAMPING * norm3(w);
    set3(F, -B_MASS[b] * vcom[0] * kl, -B_MASS[b] * vcom[1] * kl, B_MASS[b] * GRAVITY_Z - B_MASS[b] * vcom[2] * kl);
    matvec3(Iw, w, Iww);
    set3(T, -Iww[0] * ka, -Iww[1] * ka, -Iww[2] * ka);
    double fext[6], cf[3], Iv[6], vIv[6];
    cross3(com, F, cf);
    add3(fext, cf, T);
    copy3(fext + 3, F);
    mat6vec(IA[b], e->vel[b], Iv);
    crf(e->vel[b], Iv, vIv);
    for (int k = 0; k < 6; k++) pA[b][k] = vIv[k] - fext[k];
  }
  for (int b = NB - 1; b >= 0; b--) {
    double Ia[36], pa[6];
    if (B_DOF[b] >= 0) {
      mat6vec(IA[b], e->S[b], e->U[b]);
      double D = dot6(e->S[b], e->U[b]);
      e->Dinv[b] = 1.0 / D;
      u[b] = tau[B_DOF[b]] - dot6(e->S[b], pA[b]);
      for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) Ia[6 * i + j] = IA[b][6 * i + j] - e->U[b][i] * e->U[b][j] * e->Dinv[b];
      double Iac[6];
      mat6vec(Ia, c[b], Iac);
      for (int k = 0; k < 6; k++) pa[k] = pA[b][k] + Iac[k] + e->U[b][k] * (e->Dinv[b] * u[b]);
    } else {
      memcpy(Ia, IA[b], sizeof Ia);
      double Iac[6];
      mat6vec(Ia, c[b], Iac);
      for (int k = 0; k < 6; k++) pa[k] = pA[b][k] + Iac[k];
      e->Dinv[b] = 0; u[b] = 0;
      memset(e->U[b], 0, sizeof(double) * 6);
    }
    int p = B_PARENT[b];
    if (p >= 0) {
      for (int k = 0; k < 36; k++) IA[p][k] += Ia[k];
      for (int k = 0; k < 6; k++) pA[p][k] += pa[k];
    }
  }
  for (int b = 0; b < NB; b++) {
    for (int k = 0; k < 6; k++) a[b][k] = (B_PARENT[b] < 0 ? 0.0 : a[B_PARENT[b]][k]) + c[b][k];
    if (B_DOF[b] >= 0) {
      double qa = e->Dinv[b] * (u[b] - dot6(e->U[b], a[b]));
      qdd[B_DOF[b]] = qa;
      for (int k = 0; k < 6; k++) a[b][k] += e->S[b][k] * qa;
    }
  }
}

/* Unit-impulse response (btMultiBody::calcAccelerationDeltasMultiDof): generalised test force
 * tau[ND] plus spatial test forces f[NB][6] (at the world origin) -> delta qd = M^-1 J^T. */
static void impulse_response(const PmgoEnv* e, const double f[NB][6], const double* tau, double* dqd) {
  double pa[NB][6], u[NB], a[NB][6];
  for (int b = 0; b < NB; b++) for (int k = 0; k < 6; k++) pa[b][k] = f ? -f[b][k] : 0.0;
  for (int b = NB - 1; b >= 0; b--) {
    int p = B_PARENT[b];
    if (B_DOF[b] >= 0) {
      u[b] = (tau ? tau[B_DOF[b]] : 0.0) - dot6(e->S[b], pa[b]);
      if (p >= 0) for (int k = 0; k < 6; k++) pa[p][k] += pa[b][k] + e->U[b][k] * (e->Dinv[b] * u[b]);
    } else if (p >= 0) {
      for (int k = 0; k < 6; k++) pa[p][k] += pa[b][k];
    }
  }
  for (int b = 0; b < NB; b++) {
    for (int k = 0; k < 6; k++) a[b][k] = B_PARENT[b] < 0 ? 0.0 : a[B_PARENT[b]][k];
    if (B_DOF[b] >= 0) {
      double qa = e->Dinv[b] * (u[b] - dot6(e->U[b], a[b]));
      dqd[B_DOF[b]] = qa;
      for (int k = 0; k < 6; k++) a[b][k] += e->S[b][k] * qa;
    }
  }
}

/* Jacobian row of a unit force `dir` applied at world point `p` on robot body `body`, and its
 * M^-1 J^T (fillContactJacobianMultiDof + calcAccelerationDeltasMultiDof). */
static void robot_point_jacobian(const PmgoEnv* e, int body, const double* p, const double* dir, double* J, double* MJ) {
  double f[NB][6], sf[6];
  memset(f, 0, sizeof f);
  cross3(p, dir, sf);
  copy3(sf + 3, dir);
  memcpy(f[body], sf, sizeof sf);
  for (int d = 0; d < ND; d++) J[d] = 0;
  for (int b = body; b >= 0; b = B_PARENT[b])
    if (B_DOF[b] >= 0) J[B_DOF[b]] = dot6(e->S[b], sf);
  impulse_response(e, f, NULL, MJ);
}

/* ------------------------------------------------------------------------------------------ */
/* box-box contact generation (SAT + face clipping, after the ODE-derived btBoxBoxDetector)   */
/* ------------------------------------------------------------------------------------------ */
typedef struct { double pB[3], nB[3], dist; } ContactOut;

static int clip_quad_to_rect(const double h[2], const double quad[8], double out[16]) {
  double buf[2][16];
  int nq = 4, nr = 0, cur = 0;
  memcpy(buf[0], quad, sizeof(double) * 8);
  for (int dir = 0; dir <= 1; dir++) {
    for (int sign = -1; sign <= 1; sign += 2) {
      const double* q = buf[cur];
      double* r = buf[cur ^ 1];
      nr = 0;
      int full = 0;
      for (int i = 0; i < nq && !full; i++) {
        const double* pq = q + 2 * i;
        const double* nx = q + 2 * ((i + 1) % nq);
        int in0 = sign * pq[dir] < h[dir], in1 = sign * nx[dir] < h[dir];
        if (in0) {
          r[2 * nr] = pq[0]; r[2 * nr + 1] = pq[1]; nr++;
          if (nr & 8) { full = 1; break; }
        }
        if (in0 ^ in1) {
          r[2 * nr + (1 - dir)] = pq[1 - dir] + (nx[1 - dir] - pq[1 - dir]) / (nx[dir] - pq[dir]) * (sign * h[dir] - pq[dir]);
          r[2 * nr + dir] = sign * h[dir];
          nr++;
          if (nr & 8) { full = 1; break; }
        }
      }
      cur ^= 1;
      nq = nr;
      if (full) goto done;
    }
  }
done:
  memcpy(out, buf[cur], sizeof(double) * 2 * nr);
  return nr;
}

static void cull_points(int n, const double* p, int m, int i0, int* iret) {
  double a, cx, cy, q;
  if (n == 1) { cx = p[0]; cy = p[1]; }
  else if (n == 2) { cx = 0.5 * (p[0] + p[2]); cy = 0.5 * (p[1] + p[3]); }
  else {
    a = 0; cx = 0; cy = 0;
    for (int i = 0; i < n - 1; i++) {
      q = p[2 * i] * p[2 * i + 3] - p[2 * i + 2] * p[2 * i + 1];
      a += q; cx += q * (p[2 * i] + p[2 * i + 2]); cy += q * (p[2 * i + 1] + p[2 * i + 3]);
    }
    q = p[2 * n - 2] * p[1] - p[0] * p[2 * n - 1];
    a = fabs(a + q) > DBL_EPSILON ? 1.0 / (3.0 * (a + q)) : 1e18;
    cx = a * (cx + q * (p[2 * n - 2] + p[0]));
    cy = a * (cy + q * (p[2 * n - 1] + p[1]));
  }
  double A[8]; int avail[8];
  for (int i = 0; i < n; i++) { A[i] = atan2(p[2 * i + 1] - cy, p[2 * i] - cx); avail[i] = 1; }
  avail[i0] = 0; iret[0] = i0;
  for (int j = 1; j < m; j++) {
    a = j * (2 * M_PI / m) + A[i0];
    if (a > M_PI) a -= 2 * M_PI;
    double best = 1e9; int bi = i0;
    for (int i = 0; i < n; i++) if (avail[i]) {
      double d = fabs(A[i] - a);
      if (d > M_PI) d = 2 * M_PI - d;
      if (d < best) { best = d; bi = i; }
    }
    avail[bi] = 0; iret[j] = bi;
  }
}

/* boxes given by centre p, rotation R (row-major, columns = box axes) and half extents.
 * Returns up to 4 contacts: point on B, normal on B (B -> A), signed distance (<= 0). */
static int box_box(const double* p1, const double* R1, const double* A, const double* p2, const double* R2,
                   const double* B, ContactOut* out) {
  double p[3], pp[3], R[3][3], Q[3][3];
  sub3(p, p2, p1);
  matTvec3(R1, p, pp);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      R[i][j] = R1[i] * R2[j] + R1[3 + i] * R2[3 + j] + R1[6 + i] * R2[6 + j];
      Q[i][j] = fabs(R[i][j]);
    }
  double s = -DBL_MAX, s2;
  int code = 0, invert = 0;
  double normal[3] = {0, 0, 0}, normalC[3] = {0, 0, 0};
  int normal_from_R = 0; /* 1: column of R1, 2: column of R2 */
  int ncol = 0;
  for (int i = 0; i < 3; i++) { /* face axes of box 1 */
    s2 = fabs(pp[i]) - (A[i] + B[0] * Q[i][0] + B[1] * Q[i][1] + B[2] * Q[i][2]);
    if (s2 > 0) return 0;
    if (s2 > s) { s = s2; normal_from_R = 1; ncol = i; invert = pp[i] < 0; code = i + 1; }
  }
  for (int j = 0; j < 3; j++) { /* face axes of box 2 */
    double e1 = R2[j] * p[0] + R2[3 + j] * p[1] + R2[6 + j] * p[2];
    s2 = fabs(e1) - (A[0] * Q[0][j] + A[1] * Q[1][j] + A[2] * Q[2][j] + B[j]);
    if (s2 > 0) return 0;
    if (s2 > s) { s = s2; normal_from_R = 2; ncol = j; invert = e1 < 0; code = j + 4; }
  }
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Q[i][j] += 1.0e-5; /* fudge2 */
  for (int i = 0; i < 3; i++) { /* edge x edge axes */
    int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
    for (int j = 0; j < 3; j++) {
      int j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      double e1 = pp[i2] * R[i1][j] - pp[i1] * R[i2][j];
      double e2 = A[i1] * Q[i2][j] + A[i2] * Q[i1][j] + B[j1] * Q[i][j2] + B[j2] * Q[i][j1];
      double n[3];
      n[i] = 0; n[i1] = -R[i2][j]; n[i2] = R[i1][j];
      s2 = fabs(e1) - e2;
      if (s2 > DBL_EPSILON) return 0;
      double l = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
      if (l > DBL_EPSILON) {
        s2 /= l;
        if (s2 * BOX_FUDGE > s) {
          s = s2; normal_from_R = 0;
          set3(normalC, n[0] / l, n[1] / l, n[2] / l);
          invert = e1 < 0; code = 7 + 3 * i + j;
        }
      }
    }
  }
  if (!code) return 0;
  if (normal_from_R == 1) set3(normal, R1[ncol], R1[3 + ncol], R1[6 + ncol]);
  else if (normal_from_R == 2) set3(normal, R2[ncol], R2[3 + ncol], R2[6 + ncol]);
  else matvec3(R1, normalC, normal);
  if (invert) { normal[0] = -normal[0]; normal[1] = -normal[1]; normal[2] = -normal[2]; }
  double depth = -s;
  /* `normal` points from box 1 (A) to box 2 (B); Bullet reports -normal as the normal on B */
  if (code > 6) { /* edge-edge: one contact, the closest point on B's edge */
    double pa[3], pb[3], ua[3], ub[3];
    copy3(pa, p1); copy3(pb, p2);
    for (int j = 0; j < 3; j++) {
      double col[3] = {R1[j], R1[3 + j], R1[6 + j]};
      axpy3(pa, (dot3(normal, col) > 0 ? 1.0 : -1.0) * A[j], col);
      double col2[3] = {R2[j], R2[3 + j], R2[6 + j]};
      axpy3(pb, (dot3(normal, col2) > 0 ? -1.0 : 1.0) * B[j], col2);
    }
    int ia = (code - 7) / 3, ib = (code - 7) % 3;
    set3(ua, R1[ia], R1[3 + ia], R1[6 + ia]);
    set3(ub, R2[ib], R2[3 + ib], R2[6 + ib]);
    double d[3];
    sub3(d, pb, pa);
    double uaub = dot3(ua, ub), q1 = dot3(ua, d), q2 = -dot3(ub, d), den = 1 - uaub * uaub;
    double beta = den <= 1e-4 ? 0.0 : (uaub * q1 + q2) / den;
    axpy3(pb, beta, ub);
    copy3(out[0].pB, pb);
    set3(out[0].nB, -normal[0], -normal[1], -normal[2]);
    out[0].dist = -depth;
    return 1;
  }
  /* face contact: reference face on `a`, incident face on `b` */
  const double *Ra, *Rb, *pa, *pb, *Sa, *Sb;
  double normal2[3];
  if (code <= 3) { Ra = R1; Rb = R2; pa = p1; pb = p2; Sa = A; Sb = B; copy3(normal2, normal); }
  else { Ra = R2; Rb = R1; pa = p2; pb = p1; Sa = B; Sb = A; set3(normal2, -normal[0], -normal[1], -normal[2]); }
  double nr[3], anr[3];
  matTvec3(Rb, normal2, nr);
  for (int k = 0; k < 3; k++) anr[k] = fabs(nr[k]);
  int lanr, a1, a2;
  if (anr[1] > anr[0]) {
    if (anr[1] > anr[2]) { a1 = 0; lanr = 1; a2 = 2; } else { a1 = 0; a2 = 1; lanr = 2; }
  } else {
    if (anr[0] > anr[2]) { lanr = 0; a1 = 1; a2 = 2; } else { a1 = 0; a2 = 1; lanr = 2; }
  }
  double center[3];
  for (int k = 0; k < 3; k++)
    center[k] = pb[k] - pa[k] + (nr[lanr] < 0 ? 1.0 : -1.0) * Sb[lanr] * Rb[3 * k + lanr];
  int codeN = code <= 3 ? code - 1 : code - 4, code1, code2;
  if (codeN == 0) { code1 = 1; code2 = 2; } else if (codeN == 1) { code1 = 0; code2 = 2; } else { code1 = 0; code2 = 1; }
  double ra1[3] = {Ra[code1], Ra[3 + code1], Ra[6 + code1]}, ra2[3] = {Ra[code2], Ra[3 + code2], Ra[6 + code2]};
  double rb1[3] = {Rb[a1], Rb[3 + a1], Rb[6 + a1]}, rb2[3] = {Rb[a2], Rb[3 + a2], Rb[6 + a2]};
  double c1 = dot3(center, ra1), c2 = dot3(center, ra2);
  double m11 = dot3(ra1, rb1), m12 = dot3(ra1, rb2), m21 = dot3(ra2, rb1), m22 = dot3(ra2, rb2);
  double quad[8];
  {
    double k1 = m11 * Sb[a1], k2 = m21 * Sb[a1], k3 = m12 * Sb[a2], k4 = m22 * Sb[a2];
    quad[0] = c1 - k1 - k3; quad[1] = c2 - k2 - k4;
    quad[2] = c1 - k1 + k3; quad[3] = c2 - k2 + k4;
    quad[4] = c1 + k1 + k3; quad[5] = c2 + k2 + k4;
    quad[6] = c1 + k1 - k3; quad[7] = c2 + k2 - k4;
  }
  double rect[2] = {Sa[code1], Sa[code2]}, ret[16];
  int n = clip_quad_to_rect(rect, quad, ret);
  if (n < 1) return 0;
  double point[8][3], dep[8];
  double det1 = 1.0 / (m11 * m22 - m12 * m21);
  m11 *= det1; m12 *= det1; m21 *= det1; m22 *= det1;
  int cnum = 0;
  for (int j = 0; j < n; j++) {
    double k1 = m22 * (ret[2 * j] - c1) - m12 * (ret[2 * j + 1] - c2);
    double k2 = -m21 * (ret[2 * j] - c1) + m11 * (ret[2 * j + 1] - c2);
    for (int k = 0; k < 3; k++) point[cnum][k] = center[k] + k1 * rb1[k] + k2 * rb2[k];
    dep[cnum] = Sa[codeN] - dot3(normal2, point[cnum]);
    if (dep[cnum] >= 0) { ret[2 * cnum] = ret[2 * j]; ret[2 * cnum + 1] = ret[2 * j + 1]; cnum++; }
  }
  if (cnum < 1) return 0;
  int maxc = 4, idx[8];
  if (maxc > cnum) maxc = cnum;
  if (cnum <= maxc) { for (int j = 0; j < cnum; j++) idx[j] = j; }
  else {
    int i1 = 0; double maxdepth = dep[0];
    for (int i = 1; i < cnum; i++) if (dep[i] > maxdepth) { maxdepth = dep[i]; i1 = i; }
    cull_points(cnum, ret, maxc, i1, idx);
  }
  for (int j = 0; j < maxc; j++) {
    int k = idx[j];
    double w[3];
    add3(w, point[k], pa); /* incident-face point in world */
    if (code >= 4) axpy3(w, -dep[k], normal); /* incident face was on A: move onto B */
    copy3(out[j].pB, w);
    set3(out[j].nB, -normal[0], -normal[1], -normal[2]);
    out[j].dist = -dep[k];
  }
  return maxc;
}

/* ------------------------------------------------------------------------------------------ */
/* collision pairs and persistent manifolds                                                   */
/* ------------------------------------------------------------------------------------------ */
/* ------------------------------------------------------------------------------------------ */
/* box (A) against a cylinder (B, axis = its local z): the Slide puck                         */
/* ------------------------------------------------------------------------------------------ */
/* Bullet runs GJK/EPA + a persistent manifold for this pair; the single closest-point pair it returns per call on
 * a flat-on-flat contact depends on the simplex history and cannot be restated without Bullet itself.  This is a
 * MODEL of the pair (Slide): a separating-axis choice between the cylinder axis (cap against a box face)
 * and the radial direction (curved side against the box), then
 *   cap - face : the rim points of that cap at four azimuths fixed in the cylinder (+-x, +-y) plus the deepest rim
 *                point, kept when they lie over the face and penetrate it, and the corners of the face that lie
 *                inside the cap disc (a finger pressing on the puck); normal = the face normal; <= 4 points
 *                (the deepest first, then the ones that spread the patch);
 *   side       : the closest points of the box to the cylinder axis at the two ends of their common height range.
 * Output convention of box_box: point on B, normal on B (pointing from B towards A), signed distance (<= 0). */
static int box_cyl(const double* pa, const double* Ra, const double* ha, const double* pb, const double* Rb,
                   double r, double h, ContactOut* out) {
  double ax[3] = {Rb[2], Rb[5], Rb[8]};                     /* cylinder axis in world */
  double d[3], cb[3];                                       /* box centre in the cylinder frame */
  sub3(d, pa, pb); matTvec3(Rb, d, cb);
  double Aq[3][3];                                          /* box axes in the cylinder frame: Aq[k] = Rb^T Ra[:,k] */
  for (int k = 0; k < 3; k++) { double col[3] = {Ra[k], Ra[3 + k], Ra[6 + k]}; matTvec3(Rb, col, Aq[k]); }
  double ez = ha[0] * fabs(Aq[0][2]) + ha[1] * fabs(Aq[1][2]) + ha[2] * fabs(Aq[2][2]);   /* box extent along the axis */
  /* axial separation for the two caps: cap s faces the box when the box lies on its side */
  int scap = cb[2] >= 0 ? 1 : -1;
  double sep_cap = scap * cb[2] - ez - h;
  if (sep_cap > 0) return 0;
  /* radial: closest point of the box to the axis at the two ends of the common height range */
  double z0 = cb[2] - ez > -h ? cb[2] - ez : -h, z1 = cb[2] + ez < h ? cb[2] + ez : h;
  double zs[2] = {z0 + 0.05 * (z1 - z0), z1 - 0.05 * (z1 - z0)};
  double qs[2][3], rho[2];
  double sep_rad = 1e30;
  for (int i = 0; i < 2; i++) {
    double rel[3] = {-cb[0], -cb[1], zs[i] - cb[2]}, loc[3], q[3] = {cb[0], cb[1], cb[2]};
    for (int k = 0; k < 3; k++) {                           /* clamp in the box frame */
      loc[k] = rel[0] * Aq[k][0] + rel[1] * Aq[k][1] + rel[2] * Aq[k][2];
      if (loc[k] > ha[k]) loc[k] = ha[k];
      if (loc[k] < -ha[k]) loc[k] = -ha[k];
      q[0] += loc[k] * Aq[k][0]; q[1] += loc[k] * Aq[k][1]; q[2] += loc[k] * Aq[k][2];
    }
    copy3(qs[i], q);
    rho[i] = sqrt(q[0] * q[0] + q[1] * q[1]);
    if (rho[i] - r < sep_rad) sep_rad = rho[i] - r;
  }
  if (sep_rad > 0) return 0;
  int n = 0;
  if (sep_cap >= sep_rad) {
    /* ---- cap against the box face whose outward normal opposes the cap normal most ---- */
    int fj = 0; double best = -1;
    for (int k = 0; k < 3; k++) if (fabs(Aq[k][2]) > best) { best = fabs(Aq[k][2]); fj = k; }
    double sig = Aq[fj][2] * scap > 0 ? -1.0 : 1.0;         /* the face looks against the cap normal */
    double nf[3] = {sig * Aq[fj][0], sig * Aq[fj][1], sig * Aq[fj][2]};   /* outward face normal, cylinder frame */
    int k1 = (fj + 1) % 3, k2 = (fj + 2) % 3;
    double cand[9][7]; int nc = 0;                          /* point on B (cyl frame), distance, point on A */
    /* rim points of the cap */
    double u[5][2] = {{1, 0}, {-1, 0}, {0, 1}, {0, -1}, {0, 0}};
    double tx = -nf[0], ty = -nf[1], tn = sqrt(tx * tx + ty * ty);
    int nu = 4;
    if (tn > 0.02) { u[4][0] = tx / tn; u[4][1] = ty / tn; nu = 5; }   /* tilted by more than ~1 degree: the lowest rim point matters */
    for (int i = 0; i < nu; i++) {
      double p[3] = {r * u[i][0], r * u[i][1], scap * h}, rel[3], loc[3];
      sub3(rel, p, cb);
      for (int k = 0; k < 3; k++) loc[k] = rel[0] * Aq[k][0] + rel[1] * Aq[k][1] + rel[2] * Aq[k][2];
      if (fabs(loc[k1]) > ha[k1] || fabs(loc[k2]) > ha[k2]) continue;
      double dist = sig * loc[fj] - ha[fj];
      if (dist > 0) continue;
      copy3(cand[nc], p); cand[nc][3] = dist;
      nc++;
    }
    /* corners of the face inside the cap disc */
    double ncn = -(nf[2] * scap);                           /* n . cap normal with n = -nf */
    if (ncn > 1e-6)
      for (int i = 0; i < 4; i++) {
        double loc[3]; loc[fj] = sig * ha[fj]; loc[k1] = (i & 1 ? 1 : -1) * ha[k1]; loc[k2] = (i & 2 ? 1 : -1) * ha[k2];
        double v[3] = {cb[0], cb[1], cb[2]};
        for (int k = 0; k < 3; k++) { v[0] += loc[k] * Aq[k][0]; v[1] += loc[k] * Aq[k][1]; v[2] += loc[k] * Aq[k][2]; }
        double dcap = (v[2] - scap * h) * scap;             /* height of the corner above the cap plane */
        double dist = dcap / ncn;                            /* along n = -nf */
        if (dist > 0) continue;
        double pB[3] = {v[0] + dist * nf[0], v[1] + dist * nf[1], v[2] + dist * nf[2]};   /* pB = pA - dist n */
        if (pB[0] * pB[0] + pB[1] * pB[1] > r * r) continue;
        copy3(cand[nc], pB); cand[nc][3] = dist;
        nc++;
      }
    /* keep <= 4: the deepest, then the candidates farthest from the ones already kept */
    int used[9] = {0}, keep[4];
    for (int m = 0; m < 4 && m < nc; m++) {
      int bi = -1; double bv = -1e30;
      for (int i = 0; i < nc; i++) {
        if (used[i]) continue;
        double score;
        if (m == 0) score = -cand[i][3];
        else {
          score = 1e30;
          for (int j = 0; j < m; j++) {
            double e0 = cand[i][0] - cand[keep[j]][0], e1 = cand[i][1] - cand[keep[j]][1], e2 = cand[i][2] - cand[keep[j]][2];
            double dd = e0 * e0 + e1 * e1 + e2 * e2;
            if (dd < score) score = dd;
          }
          if (score < 1e-10) continue;                      /* coincides with a kept point (deepest rim point = a fixed one) */
        }
        if (score > bv) { bv = score; bi = i; }
      }
      if (bi < 0) break;
      used[bi] = 1; keep[m] = bi;
      double nw[3] = {-nf[0], -nf[1], -nf[2]}, pw[3];
      matvec3(Rb, cand[bi], pw); add3(out[n].pB, pw, pb);
      matvec3(Rb, nw, out[n].nB);
      out[n].dist = cand[bi][3];
      n++;
    }
    return n;
  }
  /* ---- curved side against the box ---- */
  for (int i = 0; i < 2; i++) {
    if (rho[i] - r > 0 || rho[i] < 1e-9) continue;
    if (i == 1 && fabs(zs[1] - zs[0]) < 1e-6) break;
    double nl[3] = {qs[i][0] / rho[i], qs[i][1] / rho[i], 0}, pl[3] = {nl[0] * r, nl[1] * r, qs[i][2]}, pw[3];
    matvec3(Rb, pl, pw); add3(out[n].pB, pw, pb);
    matvec3(Rb, nl, out[n].nB);
    out[n].dist = rho[i] - r;
    n++;
  }
  (void)ax;
  return n;
}

static double box_radius(const double* h) { return sqrt(h[0] * h[0] + h[1] * h[1] + h[2] * h[2]); }

static Geom make_geom(int kind, int index, const double* half, double friction) {
  Geom g;
  g.kind = kind; g.index = index; copy3(g.half, half); g.friction = friction; g.radius = box_radius(half); g.cylinder = 0;
  return g;
}

static void build_pairs(PmgoEnv* e) {
  double bh[3] = {PMG_BLOCK_HALF, PMG_BLOCK_HALF, PMG_BLOCK_HALF};
  Geom table = make_geom(0, 0, TABLE_HALF, PMG_TABLE_FRICTION), floor_ = make_geom(0, 1, FLOOR_HALF, PMG_FLOOR_FRICTION);
  if (e->puck) {
    table = make_geom(0, 0, LONG_TABLE_HALF, PMG_LONG_TABLE_FRICTION);
    set3(bh, PMG_PUCK_RADIUS, PMG_PUCK_RADIUS, PMG_PUCK_HALF_LEN);
  }
  Geom f1 = make_geom(1, PMG_BODY_FINGER1, FINGER_HALF, PMG_FINGER_FRICTION), f2 = make_geom(1, PMG_BODY_FINGER2, FINGER_HALF, PMG_FINGER_FRICTION);
  int n = 0;
  e->pairs[n].a = f1; e->pairs[n++].b = table;
  e->pairs[n].a = f2; e->pairs[n++].b = table;
  for (int i = 0; i < e->nb; i++) {
    Geom blk = make_geom(2, i, bh, e->puck ? PMG_PUCK_FRICTION : PMG_BLOCK_FRICTION);
    blk.cylinder = e->puck;
    e->pairs[n].a = table; e->pairs[n++].b = blk;
    e->pairs[n].a = floor_; e->pairs[n++].b = blk;
    e->pairs[n].a = f1; e->pairs[n++].b = blk;
    e->pairs[n].a = f2; e->pairs[n++].b = blk;
  }
  for (int i = 0; i < e->nb; i++)
    for (int j = i + 1; j < e->nb; j++) {
      e->pairs[n].a = make_geom(2, i, bh, PMG_BLOCK_FRICTION);
      e->pairs[n++].b = make_geom(2, j, bh, PMG_BLOCK_FRICTION);
    }
  /* Multi-block scenes: the gripper-base cylinder (r 0.05, 0.045 - 0.085 above the tip) against every block -- it
   * reaches the top of a stack of two or more blocks (hard part 4; one block on the table stays
   * 3 cm below it at the lowest tip height).  Default lateral friction (the link has no <contact> tag).  Body A is
   * the cylinder: box_cyl runs with the roles exchanged and collide() turns its contacts round. */
  if (e->nb >= 2) {
    double gh[3] = {PMG_GBASE_RADIUS, PMG_GBASE_RADIUS, PMG_GBASE_HALF_LEN};
    Geom gb = make_geom(1, PMG_BODY_GBASE, gh, PMG_GBASE_FRICTION);
    gb.cylinder = 1;
    for (int i = 0; i < e->nb; i++) {
      e->pairs[n].a = gb;
      e->pairs[n++].b = make_geom(2, i, bh, PMG_BLOCK_FRICTION);
    }
  }
  e->npair = n;
  memset(e->man, 0, sizeof e->man);
}

static const double I3c[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
static void geom_pose(const PmgoEnv* e, const Geom* g, const double** p, const double** R) {
  if (g->kind == 0) { *p = g->index == 0 ? e->table_center : FLOOR_CENTER; *R = I3c; }
  else if (g->kind == 1) { *p = e->bp[g->index]; *R = e->bR[g->index]; }
  else { *p = e->bpos[g->index]; *R = e->blkR[g->index]; }
}

static void manifold_add(Manifold* m, double thr, const double* lA, const double* lB, const double* nB, double dist) {
  /* btPersistentManifold::getCacheEntry / addManifoldPoint / replaceContactPoint / sortCachedPoints */
  double shortest = thr * thr;
  int nearest = -1;
  for (int i = 0; i < m->n; i++) {
    double d[3];
    sub3(d, m->lA[i], lA);
    double dd = dot3(d, d);
    if (dd < shortest) { shortest = dd; nearest = i; }
  }
  int slot = nearest;
  if (slot < 0) {
    if (m->n < 4) slot = m->n++;
    else {
      int deepest = -1; double maxpen = dist;
      for (int i = 0; i < 4; i++) if (m->dist[i] < maxpen) { deepest = i; maxpen = m->dist[i]; }
      double res[4] = {0, 0, 0, 0}, a[3], b[3], c[3];
      if (deepest != 0) { sub3(a, lA, m->lA[1]); sub3(b, m->lA[3], m->lA[2]); cross3(a, b, c); res[0] = dot3(c, c); }
      if (deepest != 1) { sub3(a, lA, m->lA[0]); sub3(b, m->lA[3], m->lA[2]); cross3(a, b, c); res[1] = dot3(c, c); }
      if (deepest != 2) { sub3(a, lA, m->lA[0]); sub3(b, m->lA[3], m->lA[1]); cross3(a, b, c); res[2] = dot3(c, c); }
      if (deepest != 3) { sub3(a, lA, m->lA[0]); sub3(b, m->lA[2], m->lA[1]); cross3(a, b, c); res[3] = dot3(c, c); }
      slot = 0; double best = fabs(res[0]);
      for (int i = 1; i < 4; i++) if (fabs(res[i]) > best) { best = fabs(res[i]); slot = i; }
    }
  }
  copy3(m->lA[slot], lA); copy3(m->lB[slot], lB); copy3(m->nB[slot], nB); m->dist[slot] = dist;
}

static void manifold_remove(Manifold* m, int i) {
  int last = m->n - 1;
  if (i != last) {
    copy3(m->lA[i], m->lA[last]); copy3(m->lB[i], m->lB[last]); copy3(m->nB[i], m->nB[last]); m->dist[i] = m->dist[last];
  }
  m->n--;
}

static void collide(PmgoEnv* e) {
  for (int b = 0; b < e->nb; b++) quat_to_mat(e->bquat[b], e->blkR[b]);
  for (int k = 0; k < e->npair; k++) {
    const Pair* pr = &e->pairs[k];
    Manifold* m = &e->man[k];
    const double *pa, *Ra, *pb, *Rb;
    geom_pose(e, &pr->a, &pa, &Ra);
    geom_pose(e, &pr->b, &pb, &Rb);
    /* broadphase: world AABBs, each grown by gContactBreakingThreshold */
    int overlap = 1;
    for (int ax = 0; ax < 3; ax++) {
      double ea = fabs(Ra[3 * ax]) * pr->a.half[0] + fabs(Ra[3 * ax + 1]) * pr->a.half[1] + fabs(Ra[3 * ax + 2]) * pr->a.half[2] + BROADPHASE_MARGIN;
      double eb = fabs(Rb[3 * ax]) * pr->b.half[0] + fabs(Rb[3 * ax + 1]) * pr->b.half[1] + fabs(Rb[3 * ax + 2]) * pr->b.half[2] + BROADPHASE_MARGIN;
      if (fabs(pa[ax] - pb[ax]) > ea + eb) overlap = 0;
    }
    if (!overlap) { m->n = 0; continue; }
    double thr = BREAKING_THRESHOLD_FACTOR * (pr->a.radius < pr->b.radius ? pr->a.radius : pr->b.radius);
    ContactOut c[4];
    int nc;
    if (pr->a.cylinder) { /* cylinder A against box B: box_cyl with the roles exchanged, contacts turned round */
      nc = box_cyl(pb, Rb, pr->b.half, pa, Ra, pr->a.half[0], pr->a.half[2], c);
      for (int i = 0; i < nc; i++) { /* point on the box = point on the cylinder + n * distance; the normal flips */
        axpy3(c[i].pB, c[i].dist, c[i].nB);
        for (int k = 0; k < 3; k++) c[i].nB[k] = -c[i].nB[k];
      }
    } else
      nc = pr->b.cylinder ? box_cyl(pa, Ra, pr->a.half, pb, Rb, pr->b.half[0], pr->b.half[2], c)
                          : box_box(pa, Ra, pr->a.half, pb, Rb, pr->b.half, c);
    for (int i = 0; i < nc; i++) {
      double wa[3], lA[3], lB[3], t[3];
      copy3(wa, c[i].pB);
      axpy3(wa, c[i].dist, c[i].nB); /* point on A = point on B + n * distance */
      sub3(t, wa, pa); matTvec3(Ra, t, lA);
      sub3(t, c[i].pB, pb); matTvec3(Rb, t, lB);
      manifold_add(m, thr, lA, lB, c[i].nB, c[i].dist);
    }
    /* refreshContactPoints */
    for (int i = m->n - 1; i >= 0; i--) {
      double wa[3], wb[3], d[3];
      matvec3(Ra, m->lA[i], wa); add3(wa, wa, pa);
      matvec3(Rb, m->lB[i], wb); add3(wb, wb, pb);
      sub3(d, wa, wb);
      m->dist[i] = dot3(d, m->nB[i]);
      if (m->dist[i] > thr) { manifold_remove(m, i); continue; }
      double proj[3], pd[3];
      copy3(proj, wa); axpy3(proj, -m->dist[i], m->nB[i]);
      sub3(pd, wb, proj);
      if (dot3(pd, pd) > thr * thr) manifold_remove(m, i);
    }
  }
}

/* ------------------------------------------------------------------------------------------ */
/* constraint rows + projected Gauss-Seidel                                                   */
/* ------------------------------------------------------------------------------------------ */
static void plane_space(const double* n, double* p, double* q) { /* btPlaneSpace1 */
  if (fabs(n[2]) > 0.7071067811865475244008443621048490) {
    double a = n[1] * n[1] + n[2] * n[2], k = 1.0 / sqrt(a);
    p[0] = 0; p[1] = -n[2] * k; p[2] = n[1] * k;
    q[0] = a * k; q[1] = -n[0] * p[2]; q[2] = n[0] * p[1];
  } else {
    double a = n[0] * n[0] + n[1] * n[1], k = 1.0 / sqrt(a);
    p[0] = -n[1] * k; p[1] = n[0] * k; p[2] = 0;
    q[0] = -n[2] * p[1]; q[1] = n[2] * p[0]; q[2] = a * k;
  }
}

static void block_jac(const PmgoEnv* e, int blk, const double* pw, const double* dir, double sign, double* J, double* MJ) {
  double r[3], rxn[3];
  sub3(r, pw, e->bpos[blk]);
  cross3(r, dir, rxn);
  for (int k = 0; k < 3; k++) { J[k] = sign * rxn[k]; J[3 + k] = sign * dir[k]; }
  /* world inverse inertia R diag(1/I) R^T; the blocks are cubes so this is isotropic */
  double t[3], u[3];
  matTvec3(e->blkR[blk], J, t);
  for (int k = 0; k < 3; k++) t[k] /= e->blk_inertia[k];
  matvec3(e->blkR[blk], t, u);
  for (int k = 0; k < 3; k++) { MJ[k] = u[k]; MJ[3 + k] = J[3 + k] / e->blk_mass; }
}

static double row_velocity(const PmgoEnv* e, const Row* r, const double* qd, double bvel[MAXBLK][6]) {
  double v = 0;
  if (r->has_robot) for (int d = 0; d < ND; d++) v += r->Jr[d] * qd[d];
  if (r->blkA >= 0) v += dot6(r->JA, bvel[r->blkA]);
  if (r->blkB >= 0) v += dot6(r->JB, bvel[r->blkB]);
  (void)e;
  return v;
}

/* Set up one contact-space row (normal or friction) along `dir` (points from B to A). */
static void contact_row(PmgoEnv* e, Row* r, const Pair* pr, const double* pA, const double* pB, const double* dir) {
  memset(r, 0, sizeof *r);
  r->blkA = r->blkB = -1;
  double denom = 0;
  if (pr->a.kind == 1) {
    r->has_robot = 1;
    robot_point_jacobian(e, pr->a.index, pA, dir, r->Jr, r->MJr);
    for (int d = 0; d < ND; d++) denom += r->Jr[d] * r->MJr[d];
  } else if (pr->a.kind == 2) {
    r->blkA = pr->a.index;
    block_jac(e, r->blkA, pA, dir, 1.0, r->JA, r->MJA);
    denom += dot6(r->JA, r->MJA);
  }
  if (pr->b.kind == 2) {
    r->blkB = pr->b.index;
    block_jac(e, r->blkB, pB, dir, -1.0, r->JB, r->MJB);
    denom += dot6(r->JB, r->MJB);
  }
  r->diag_inv = 1.0 / denom;
}

static void solve_constraints(PmgoEnv* e) {
  Row* rows = e->rows;
  int nnc = 0; /* non-contact rows */
  double bvel[MAXBLK][6];
  for (int b = 0; b < e->nb; b++) { copy3(bvel[b], e->bw[b]); copy3(bvel[b] + 3, e->bv[b]); }
  /* --- joint limits and motors, in Bullet's (sorted) constraint order ----------------------- */
  for (int oi = 0; oi < 2 * ND; oi++) {
    int id = NONCONTACT_ORDER[oi];
    int d = id % ND;
    double tau[ND], col[ND];
    memset(tau, 0, sizeof tau);
    if (id < ND) { /* btMultiBodyJointLimitConstraint: row 0 lower, row 1 upper */
      for (int side = 0; side < 2; side++) {
        double pen = side == 0 ? e->q[d] - DOF_LOWER[d] : DOF_UPPER[d] - e->q[d];
        if (pen > 0) continue;
        double dir = side ? -1.0 : 1.0;
        Row* r = &rows[nnc++];
        memset(r, 0, sizeof *r);
        r->blkA = r->blkB = -1; r->has_robot = 1;
        tau[d] = dir;
        impulse_response(e, NULL, tau, col);
        r->Jr[d] = dir;
        memcpy(r->MJr, col, sizeof col);
        r->diag_inv = 1.0 / (dir * col[d]);
        double rel_vel = dir * e->qd[d];
        double pos_err = pen > SPLIT_IMPULSE_PEN_THRESHOLD ? -pen * CONTACT_ERP / DT : 0.0;
        r->rhs = (pos_err - rel_vel) * r->diag_inv;
        r->lo = 0; r->hi = LIMIT_MAX_IMPULSE;
      }
    } else { /* btMultiBodyJointMotor, POSITION_CONTROL (kuka.py:282-301) */
      Row* r = &rows[nnc++];
      memset(r, 0, sizeof *r);
      r->blkA = r->blkB = -1; r->has_robot = 1;
      tau[d] = 1.0;
      impulse_response(e, NULL, tau, col);
      r->Jr[d] = 1.0;
      memcpy(r->MJr, col, sizeof col);
      r->diag_inv = 1.0 / col[d];
      double target_vel = MOTOR_KP * (e->mot_target[d] - e->q[d]) / DT + e->qd[d] + MOTOR_KD * (0.0 - e->qd[d]);
      r->rhs = (target_vel - e->qd[d]) * r->diag_inv;
      r->lo = -e->mot_maximp[d]; r->hi = e->mot_maximp[d];
    }
  }
  /* --- contacts: one normal row + two friction rows per cached point ------------------------ */
  int nn = 0, nf = 0;
  Row* nrows = rows + nnc;
  int total_pts = 0;
  for (int k = 0; k < e->npair; k++) total_pts += e->man[k].n;
  Row* frows = nrows + total_pts;
  for (int k = 0; k < e->npair; k++) {
    const Pair* pr = &e->pairs[k];
    const Manifold* m = &e->man[k];
    const double *pa, *Ra, *pb, *Rb;
    geom_pose(e, &pr->a, &pa, &Ra);
    geom_pose(e, &pr->b, &pb, &Rb);
    for (int i = 0; i < m->n; i++) {
      double wa[3], wb[3], t1[3], t2[3];
      matvec3(Ra, m->lA[i], wa); add3(wa, wa, pa);
      matvec3(Rb, m->lB[i], wb); add3(wb, wb, pb);
      Row* r = &nrows[nn];
      contact_row(e, r, pr, wa, wb, m->nB[i]);
      double rel_vel = row_velocity(e, r, e->qd, bvel);
      double pen = m->dist[i] + LINEAR_SLOP;
      double pos_err = 0, vel_err = -rel_vel;
      if (pen > 0) vel_err -= pen / DT; else pos_err = -pen * CONTACT_ERP / DT;
      r->rhs = (pos_err + vel_err) * r->diag_inv;
      r->lo = 0; r->hi = 1e10;
      r->mu = pr->a.friction * pr->b.friction;
      plane_space(m->nB[i], t1, t2);
      for (int f = 0; f < 2; f++) {
        Row* fr = &frows[nf++];
        contact_row(e, fr, pr, wa, wb, f ? t2 : t1);
        fr->rhs = -row_velocity(e, fr, e->qd, bvel) * fr->diag_inv;
        fr->normal_row = nn; fr->mu = r->mu;
      }
      nn++;
    }
  }
  /* --- PGS ---------------------------------------------------------------------------------- */
  double dqd[ND], dbv[MAXBLK][6];
  memset(dqd, 0, sizeof dqd);
  memset(dbv, 0, sizeof dbv);
#define ROW_DELTA(r) ((r)->rhs - row_velocity(e, (r), dqd, dbv) * (r)->diag_inv)
#define ROW_APPLY(r, dl) do { \
    if ((r)->has_robot) for (int d_ = 0; d_ < ND; d_++) dqd[d_] += (r)->MJr[d_] * (dl); \
    if ((r)->blkA >= 0) for (int k_ = 0; k_ < 6; k_++) dbv[(r)->blkA][k_] += (r)->MJA[k_] * (dl); \
    if ((r)->blkB >= 0) for (int k_ = 0; k_ < 6; k_++) dbv[(r)->blkB][k_] += (r)->MJB[k_] * (dl); } while (0)
  for (int it = 0; it < SOLVER_ITERS; it++) {
    double residual = 0;
    for (int j = 0; j < nnc + nn; j++) {
      /* non-contact rows run backwards on even iterations, forwards on odd ones */
      Row* r = j < nnc ? &rows[(it & 1) ? j : nnc - 1 - j] : &nrows[j - nnc];
      double dl = ROW_DELTA(r), sum = r->applied + dl;
      if (sum < r->lo) { dl = r->lo - r->applied; r->applied = r->lo; }
      else if (sum > r->hi) { dl = r->hi - r->applied; r->applied = r->hi; }
      else r->applied = sum;
      ROW_APPLY(r, dl);
      double res = dl / r->diag_inv;
      if (res * res > residual) residual = res * res;
    }
    for (int j = 0; j + 1 < nf; j += 2) { /* implicit friction cone: the two tangent rows together */
      Row *ra = &frows[j], *rb = &frows[j + 1];
      double total = nrows[ra->normal_row].applied;
      if (!(total > 0)) continue;
      double lim = ra->mu * total;
      double dA = ROW_DELTA(ra), dB = ROW_DELTA(rb);
      double sA = ra->applied + dA, sB = rb->applied + dB;
      if (sA * sA + sB * sB >= lim * lim) {
        double ang = atan2(sA, sB);
        double cA = fabs(lim * sin(ang)), cB = fabs(lim * cos(ang));
        if (sA < -cA) { dA = -cA - ra->applied; ra->applied = -cA; }
        else if (sA > cA) { dA = cA - ra->applied; ra->applied = cA; }
        else ra->applied = sA;
        if (sB < -cB) { dB = -cB - rb->applied; rb->applied = -cB; }
        else if (sB > cB) { dB = cB - rb->applied; rb->applied = cB; }
        else rb->applied = sB;
      } else { ra->applied = sA; rb->applied = sB; }
      ROW_APPLY(ra, dA);
      ROW_APPLY(rb, dB);
      double r1 = dA / ra->diag_inv, r2 = dB / rb->diag_inv;
      if (r1 * r1 > residual) residual = r1 * r1;
      if (r2 * r2 > residual) residual = r2 * r2;
    }
    if (residual <= RESIDUAL_THRESHOLD) break;
  }
#undef ROW_DELTA
#undef ROW_APPLY
  for (int d = 0; d < ND; d++) {
    e->qd[d] += dqd[d];
    if (e->qd[d] > MAX_COORD_VEL) e->qd[d] = MAX_COORD_VEL;
    if (e->qd[d] < -MAX_COORD_VEL) e->qd[d] = -MAX_COORD_VEL;
  }
  for (int b = 0; b < e->nb; b++) for (int k = 0; k < 3; k++) { e->bw[b][k] += dbv[b][k]; e->bv[b][k] += dbv[b][3 + k]; }
}

/* ------------------------------------------------------------------------------------------ */
/* one 2 ms substep (btMultiBodyDynamicsWorld::internalSingleStepSimulation)                  */
/* ------------------------------------------------------------------------------------------ */
static void substep(PmgoEnv* e) {
  double qdd[ND];
  /* forward dynamics first refreshes the link frames used by the collision pass */
  fk_bodies(e->q, e->bR, e->bp, e->S);
  collide(e);
  aba(e, e->joint_damp_tau, qdd);
  for (int d = 0; d < ND; d++) {
    e->qd[d] += qdd[d] * DT;
    if (e->qd[d] > MAX_COORD_VEL) e->qd[d] = MAX_COORD_VEL;
    if (e->qd[d] < -MAX_COORD_VEL) e->qd[d] = -MAX_COORD_VEL;
  }
  for (int b = 0; b < e->nb; b++) { /* free bodies: gravity + damping; cube => no gyroscopic term */
    if (e->puck) { /* anisotropic inertia: alpha = -I^-1 (w x I w), I = R diag(I) R^T (btMultiBody's gyroscopic term) */
      double wl[3], Iw[3], g[3], gw[3];
      matTvec3(e->blkR[b], e->bw[b], wl);
      for (int k = 0; k < 3; k++) Iw[k] = e->blk_inertia[k] * wl[k];
      cross3(wl, Iw, g);
      for (int k = 0; k < 3; k++) g[k] /= e->blk_inertia[k];
      matvec3(e->blkR[b], g, gw);
      for (int k = 0; k < 3; k++) e->bw[b][k] -= gw[k] * DT;
    }
    double kl = LINK_DAMPING + LINK_DAMPING * norm3(e->bv[b]), ka = LINK_DAMPING + LINK_DAMPING * norm3(e->bw[b]);
    for (int k = 0; k < 3; k++) {
      double acc = -e->bv[b][k] * kl + (k == 2 ? GRAVITY_Z : 0.0);
      e->bv[b][k] += acc * DT;
      e->bw[b][k] += -e->bw[b][k] * ka * DT;
    }
  }
  solve_constraints(e);
  for (int d = 0; d < ND; d++) e->q[d] += e->qd[d] * DT;
  for (int b = 0; b < e->nb; b++) {
    axpy3(e->bpos[b], DT, e->bv[b]);
    double w = norm3(e->bw[b]), ax[3];
    if (w * DT > 0.5 * M_PI * 0.5) w = 0.5 * (0.5 * M_PI) / DT; /* ANGULAR_MOTION_THRESHOLD */
    double s = w < 0.001 ? 0.5 * DT - DT * DT * DT * 0.020833333333 * w * w : sin(0.5 * w * DT) / w;
    set3(ax, e->bw[b][0] * s, e->bw[b][1] * s, e->bw[b][2] * s);
    double dq[4] = {ax[0], ax[1], ax[2], cos(0.5 * w * DT)}, nq[4];
    quat_mul(dq, e->bquat[b], nq);
    double n = sqrt(nq[0] * nq[0] + nq[1] * nq[1] + nq[2] * nq[2] + nq[3] * nq[3]);
    for (int k = 0; k < 4; k++) e->bquat[b][k] = nq[k] / n;
  }
}

void pmgo_substeps(PmgoEnv* e, int n) { for (int i = 0; i < n; i++) substep(e); }

void pmgo_step_simulation(PmgoEnv* e) {
  /* PyBullet adds the URDF joint-damping torque -damping*qd once per stepSimulation call; the
   * multibody forces are cleared after the last substep, so it acts in all 20 [BULLET-MEMORY] */
  for (int d = 0; d < ND; d++) e->joint_damp_tau[d] = -DOF_DAMPING[d] * e->qd[d];
  pmgo_substeps(e, SUBSTEPS_PER_CALL);
  memset(e->joint_damp_tau, 0, sizeof e->joint_damp_tau);
}

/* ------------------------------------------------------------------------------------------ */
/* environment plumbing                                                                       */
/* ------------------------------------------------------------------------------------------ */
static void robot_reset(PmgoEnv* e);

PmgoEnv* pmgo_create(int task, int num_block, int binary_reward, double thr, int max_steps) {
  return pmgo_create_ex(task, num_block, binary_reward, thr, max_steps, 0, 0);
}

PmgoEnv* pmgo_create_ex(int task, int num_block, int binary_reward, double thr, int max_steps,
                        int grip_informed_goal, int joint_control) {
  return pmgo_create_ex2(task, num_block, binary_reward, thr, max_steps, grip_informed_goal, joint_control, 0);
}

void pmgo_set_sub_goal(PmgoEnv* e, int ind) { if (e->td) e->sub_goal_ind = ind; }
static void write_obs(PmgoEnv* e, double* o);
void pmgo_observe(PmgoEnv* e, double* obs_out) { write_obs(e, obs_out); }

PmgoEnv* pmgo_create_ex2(int task, int num_block, int binary_reward, double thr, int max_steps,
                         int grip_informed_goal, int joint_control, int task_decomposition) {
  return pmgo_create_ex3(task, num_block, binary_reward, thr, max_steps, grip_informed_goal, joint_control, task_decomposition, 0, 0);
}

void pmgo_set_curriculum_update(PmgoEnv* e, int on) { if (e->cur) e->cur_update = on != 0; }
int pmgo_get_curriculum(const PmgoEnv* e, double* prob_out) {
  for (int k = 0; k < e->nb; k++) prob_out[k] = e->cur_prob[k];
  return e->cur_level;
}

/* kuka_multi_step_base_env.py:350-379, statement by statement (numpy negative indexing included) */
static void update_curriculum_prob(PmgoEnv* e) {
  const int n = e->nb;
  int fin[MAXBLK] = {0}, half[MAXBLK] = {0};
  if (n < 2) return; /* the reference raises IndexError on mask_finished[-2] here; pmg_create rejects the combination */
  for (int i = 0; i < n; i++) {
    fin[i] = e->cur_count[i] >= e->cur_goals_per;
    half[i] = e->cur_count[i] >= e->cur_goals_per / 2;
    if (fin[i]) e->cur_prob[i] = 0.0;
  }
  if (half[0] && !fin[0]) { e->cur_prob[0] = 0.5; e->cur_prob[1] = 0.5; }
  for (int i = 1; i < n - 1; i++)
    if (fin[i - 1] && !fin[i]) {
      if (half[i]) { e->cur_prob[i] = 0.5; e->cur_prob[i + 1] = 0.5; }
      else e->cur_prob[i] = 1.0;
    }
  if (fin[n - 2]) e->cur_prob[n - 1] = 1.0;
}

PmgoEnv* pmgo_create_ex3(int task, int num_block, int binary_reward, double thr, int max_steps,
                         int grip_informed_goal, int joint_control, int task_decomposition,
                         int use_curriculum, long num_goals_to_generate) {
  PmgoEnv* e = (PmgoEnv*)calloc(1, sizeof *e);
  e->td = task_decomposition && task == PMGO_BLOCK_STACK;
  e->sub_goal_ind = -1;
  e->cur = use_curriculum && (task == PMGO_BLOCK_STACK || task == PMGO_BLOCK_REARRANGE) && !e->td; /* mutually exclusive (:113,121) */
  if (e->cur) {
    e->cur_prob[0] = 1.0;
    e->cur_goals_per = (double)(num_goals_to_generate / num_block); /* floor division (:138) */
  }
  e->task = task; e->binary = binary_reward; e->thr = thr; e->max_steps = max_steps;
  e->has_obj = task != PMGO_REACH;
  e->multi = task == PMGO_BLOCK_STACK || task == PMGO_BLOCK_REARRANGE;
  /* kuka_multi_step_envs.py:30 (stack: grasping) / :170 (rearrange: grasping=False, start on the table) */
  e->grasping = task == PMGO_PICK_AND_PLACE || task == PMGO_BLOCK_STACK;
  e->grip = grip_informed_goal && task == PMGO_BLOCK_STACK; /* rearrange asserts it off (kuka_multi_step_envs.py:158) */
  e->jc = joint_control != 0;
  e->nb = task == PMGO_REACH ? 0 : (e->multi ? num_block : 1);
  /* kuka.py:104-118: joint control takes 7 joint deltas (+ the grip command) */
  e->adim = e->jc ? (e->grasping ? 8 : 7) : (e->grasping ? 4 : 3);
  e->target_in_air = task != PMGO_PUSH && task != PMGO_SLIDE;
  /* Slide (kuka_single_step_envs.py:49-59): long low-friction table, a 2 kg puck instead of the cube */
  e->puck = task == PMGO_SLIDE;
  e->blk_mass = e->puck ? PMG_PUCK_MASS : PMG_BLOCK_MASS;
  for (int k = 0; k < 3; k++) e->blk_inertia[k] = e->puck ? PUCK_INERTIA[k] : PMG_BLOCK_INERTIA;
  e->blk_spawn_z = e->puck ? PMG_PUCK_SPAWN_Z : BLOCK_SPAWN_Z;
  copy3(e->table_center, e->puck ? LONG_TABLE_CENTER : TABLE_CENTER);
  if (task == PMGO_REACH) { e->dims[0] = 3; e->dims[1] = 3; e->dims[2] = 3; e->dims[3] = 3; }
  else if (e->multi) { e->dims[0] = 8 + 16 * e->nb; e->dims[1] = 4 + 3 * e->nb; e->dims[2] = e->dims[3] = 3 * e->nb; }
  else { e->dims[0] = 20; e->dims[1] = 7; e->dims[2] = 3; e->dims[3] = 3; }
  if (e->grip) { e->dims[2] += 4; e->dims[3] += 4; } /* + gripper xyz + finger closeness (kuka_multi_step_base_env.py:300-304) */
  if (e->jc) { e->dims[0] += 7; e->dims[1] += 7; }   /* joint poses are prepended (kuka_single_step_base_env.py:214-216) */
  /* kuka.py:35-51 with each task's ctor args (obj_range = target_range = 0.15) */
  set3(e->tip_init, -0.52, 0.0, 0.25);
  if (task == PMGO_PUSH || task == PMGO_BLOCK_REARRANGE || task == PMGO_SLIDE) e->tip_init[2] = 0.175 + 0.001;
  const double obj_range = task == PMGO_SLIDE ? 0.1 : 0.15, target_range = task == PMGO_SLIDE ? 0.2 : 0.15;
  for (int k = 0; k < 3; k++) {
    e->obj_lo[k] = e->tip_init[k] - obj_range; e->obj_hi[k] = e->tip_init[k] + obj_range;
    e->tgt_lo[k] = e->tip_init[k] - target_range; e->tgt_hi[k] = e->tip_init[k] + target_range;
  }
  e->obj_lo[0] += 0.03; e->obj_hi[0] -= 0.03;
  e->tgt_lo[0] += 0.03; e->tgt_lo[2] = EE_LOWER[2]; e->tgt_hi[0] -= 0.03;
  if (task == PMGO_SLIDE) { e->tgt_lo[0] -= 0.4; e->tgt_hi[0] -= 0.4; } /* kuka_single_step_base_env.py:66-69 */
  memcpy(e->rest_pose, KUKA_REST_POSE, sizeof e->rest_pose);
  for (int b = 0; b < MAXBLK; b++) { e->bquat[b][3] = 1.0; }
  build_pairs(e);
  uint32_t key0 = 0;
  mt_init_by_array(&e->rng, &key0, 1);
  /* BaseBulletMGEnv.__init__ resets the robot once on its own (base_env.py:41) before its first
   * self.reset() (base_env.py:84): the rest pose gets one extra IK refinement, no RNG draw. */
  robot_reset(e);
  return e;
}
void pmgo_destroy(PmgoEnv* e) { free(e); }
int pmgo_dims(const PmgoEnv* e, int dims[4]) { memcpy(dims, e->dims, sizeof e->dims); return e->adim; }
void pmgo_seed_array(PmgoEnv* e, const uint32_t* key, int len) { mt_init_by_array(&e->rng, key, len); }
void pmgo_rng_uniform(PmgoEnv* e, double lo, double hi, int n, double* out) { for (int i = 0; i < n; i++) out[i] = mt_uniform(&e->rng, lo, hi); }
void pmgo_rng_shuffle(PmgoEnv* e, int64_t* arr, int n) {
  for (int i = n - 1; i > 0; i--) { int j = (int)mt_interval(&e->rng, (uint32_t)i); int64_t t = arr[i]; arr[i] = arr[j]; arr[j] = t; }
}

void pmgo_link_state(const PmgoEnv* ec, int which, double out[13]) {
  PmgoEnv* e = (PmgoEnv*)ec;
  fk_bodies(e->q, e->bR, e->bp, e->S);
  double vel[NB][6];
  for (int b = 0; b < NB; b++)
    for (int k = 0; k < 6; k++)
      vel[b][k] = (B_PARENT[b] < 0 ? 0.0 : vel[B_PARENT[b]][k]) + (B_DOF[b] >= 0 ? e->S[b][k] * e->qd[B_DOF[b]] : 0.0);
  int body; const double* off; static const double zero[3] = {0, 0, 0};
  switch (which) {
    case 0: body = PMG_BODY_LINK7; off = TIP_OFFSET; break;
    case 1: body = PMG_BODY_GBASE; off = zero; break;
    case 2: body = PMG_BODY_FINGER1; off = TAB1_OFFSET; break;
    case 3: body = PMG_BODY_FINGER2; off = TAB2_OFFSET; break;
    case 4: body = PMG_BODY_FINGER1; off = zero; break;
    default: body = PMG_BODY_FINGER2; off = zero; break;
  }
  double t[3], lin[3];
  matvec3(e->bR[body], off, t);
  add3(out, e->bp[body], t);
  mat_to_quat(e->bR[body], out + 3);
  cross3(vel[body], out, lin); /* v = v_O + w x p */
  add3(out + 7, lin, vel[body] + 3);
  copy3(out + 10, vel[body]);
}

static void write_obs(PmgoEnv* e, double* o) {
  /* kuka.py:227-256 + kuka_single_step_base_env.py:193-221 / kuka_multi_step_base_env.py:255-320 */
  double tip[13], base[13], tab1[13], tab2[13];
  pmgo_link_state(e, 0, tip);
  double closeness = 0.0, finger_vel = 0.0;
  if (e->grasping) {
    pmgo_link_state(e, 1, base); pmgo_link_state(e, 2, tab1); pmgo_link_state(e, 3, tab2);
    double d[3];
    sub3(d, tab1, tab2);
    closeness = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    finger_vel = base[7 + 1] - tab1[7 + 1];
  }
  const int jo = e->jc ? 7 : 0;
  double* obs = o + jo; double* pol = o + e->dims[0] + jo; double* ag = o + e->dims[0] + e->dims[1]; double* dg = ag + e->dims[2];
  if (e->jc) { memcpy(o, e->q, sizeof(double) * 7); memcpy(o + e->dims[0], e->q, sizeof(double) * 7); }
  if (e->task == PMGO_REACH) {
    copy3(obs, tip); copy3(pol, tip); copy3(ag, tip);
  } else if (!e->multi) {
    const double* bx = e->bpos[0];
    copy3(obs, tip); copy3(obs + 3, bx); obs[6] = closeness;
    sub3(obs + 7, tip, bx);
    copy3(obs + 10, tip + 7); obs[13] = finger_vel;
    sub3(obs + 14, tip + 7, e->bv[0]);
    sub3(obs + 17, tip + 10, e->bw[0]);
    copy3(pol, tip); pol[3] = closeness; sub3(pol + 4, tip, bx);
    copy3(ag, bx);
  } else {
    copy3(obs, tip); obs[3] = closeness; copy3(obs + 4, tip + 7); obs[7] = finger_vel;
    copy3(pol, tip); pol[3] = closeness;
    for (int n = 0; n < e->nb; n++) {
      double* bs = obs + 8 + 16 * n;
      copy3(bs, e->bpos[n]);
      sub3(bs + 3, tip, e->bpos[n]);
      memcpy(bs + 6, e->bquat[n], sizeof(double) * 4); /* "block_rpy" is the xyzw quaternion */
      sub3(bs + 10, tip + 7, e->bv[n]);
      sub3(bs + 13, tip + 10, e->bw[n]);
      sub3(pol + 4 + 3 * n, tip, e->bpos[n]);
      copy3(ag + 3 * n, e->bpos[n]);
    }
    if (e->grip) { copy3(ag + 3 * e->nb, tip); ag[3 * e->nb + 3] = closeness; }
    /* np.clip over the concatenated state / policy_state, joint poses included (:306-307) */
    for (int i = 0; i < e->dims[0]; i++) o[i] = o[i] < -5 ? -5 : (o[i] > 5 ? 5 : o[i]);
    for (int i = 0; i < e->dims[1]; i++) o[e->dims[0] + i] = o[e->dims[0] + i] < -5 ? -5 : (o[e->dims[0] + i] > 5 ? 5 : o[e->dims[0] + i]);
    /* _generate_goal(new_target=False): rebuild desired_goal from the cached order/targets */
    if (e->task == PMGO_BLOCK_STACK) {
      for (int k = 0; k < e->nb; k++) copy3(e->goal + 3 * e->last_order[k], e->last_targets[k]);
      if (e->grip) { copy3(e->goal + 3 * e->nb, e->last_targets[e->nb - 1]); e->goal[3 * e->nb + 3] = 0.03; }
    }
  }
  memcpy(dg, e->goal, sizeof(double) * e->dims[3]);
  if (e->cur && e->task == PMGO_BLOCK_REARRANGE) {
    /* kuka_multi_step_envs.py:218-225: the blocks picked for this episode (e->sub_goal_ind holds them as a bit
     * mask) take the sampled targets in order, every other block's goal is wherever it is now */
    int j = 0;
    for (int i = 0; i < e->nb; i++)
      copy3(dg + 3 * i, ((e->sub_goal_ind >> i) & 1) ? e->last_targets[j++] : e->bpos[i]);
  } else if (e->td || e->cur) {
    /* The curriculum goal of level L (kuka_multi_step_envs.py:124-148) is the "place" sub-goal of level L;
     * e->sub_goal_ind holds the equivalent sub-goal index in that mode. */
    /* kuka_multi_step_envs.py:88-120 + kuka_multi_step_base_env.py:159-165,311-313: the desired goal is
     * sub_goals[sub_goal_ind], rebuilt from the current block positions every observation.  Without the grip
     * goal sub-goal k puts the blocks of stack levels <= k on their targets and leaves the others where they
     * are; with it there is a pick (2k) and a place (2k+1) sub-goal per level and the gripper entries name the
     * block to pick / the target to place it on. */
    const int nsub = e->grip ? 2 * e->nb : e->nb;
    int ind = e->sub_goal_ind < 0 ? e->sub_goal_ind + nsub : e->sub_goal_ind; /* python list indexing */
    const int k = e->grip ? ind >> 1 : ind, place = e->grip ? (ind & 1) : 1;
    for (int i = 0; i < e->nb; i++) {
      const int b = e->last_order[i];
      const int at_target = place ? i <= k : i < k;
      copy3(dg + 3 * b, at_target ? e->last_targets[i] : e->bpos[b]);
    }
    if (e->grip) {
      copy3(dg + 3 * e->nb, place ? e->last_targets[k] : e->bpos[e->last_order[k]]);
      dg[3 * e->nb + 3] = 0.03;
    }
  }
}

static void set_arm(PmgoEnv* e, const double* pose) {
  /* Joint.reset_position: resetJointState(q, 0) + zero-force motor (robot_bases.py:230-238) */
  for (int d = 0; d < 7; d++) { e->q[d] = pose[d]; e->qd[d] = 0; e->mot_maximp[d] = 0; e->mot_target[d] = 0; }
}

static void robot_reset(PmgoEnv* e) { /* kuka.py:157-165 */
  set_arm(e, e->rest_pose);
  double qik[ND];
  pmgo_ik(e->q, e->tip_init, EE_FIXED_QUAT, 40, 1e-5, qik);
  memcpy(e->rest_pose, qik, sizeof e->rest_pose);
  set_arm(e, e->rest_pose);
  for (int d = 7; d < ND; d++) {
    e->q[d] = GRIPPER_ABS_LIMIT; e->qd[d] = 0;
    e->mot_target[d] = GRIPPER_ABS_LIMIT; e->mot_maximp[d] = FINGER_FORCE * OUTER_DT;
  }
  double quat[4];
  pmgo_fk_tip(e->q, e->ee_target, quat);
  /* kuka.py:165: joint_state_target = current joint state.  It is kept in mot_target (the motor is off until
   * the first move_arm, so the value has no effect before it is used). */
  if (e->jc) for (int d = 0; d < 7; d++) e->mot_target[d] = e->q[d];
}

static void place_blocks(PmgoEnv* e, const double* xy) {
  for (int b = 0; b < e->nb; b++) {
    set3(e->bpos[b], xy[2 * b], xy[2 * b + 1], e->blk_spawn_z);
    e->bquat[b][0] = e->bquat[b][1] = e->bquat[b][2] = 0; e->bquat[b][3] = 1;
    set3(e->bv[b], 0, 0, 0); set3(e->bw[b], 0, 0, 0);
  }
}

/* level = np_random.choice(num_curriculum, p=curriculum_prob), i.e. cdf.searchsorted(random_sample(),
 * side='right') in numpy's legacy RandomState (kuka_multi_step_envs.py:127,197) */
static int draw_curriculum_level(PmgoEnv* e) {
  double cdf[MAXBLK], acc = 0;
  for (int k = 0; k < e->nb; k++) { acc += e->cur_prob[k]; cdf[k] = acc; }
  for (int k = 0; k < e->nb; k++) cdf[k] /= acc;
  const double u = mt_double(&e->rng);
  int level = 0;
  while (level < e->nb && cdf[level] <= u) level++;
  e->cur_level = level;
  return level;
}

void pmgo_reset(PmgoEnv* e, double* obs_out) {
  robot_reset(e);
  e->elapsed = 0;
  if (!e->cur) e->sub_goal_ind = -1; /* kuka_multi_step_base_env.py:247-248 */
  double xy[2 * MAXBLK] = {0};
  if (e->multi) {
    /* kuka_multi_step_base_env.py:223-240 */
    for (int b = 0; b < e->nb; b++) {
      for (;;) {
        double x = mt_uniform(&e->rng, e->obj_lo[0], e->obj_hi[0]);
        double y = mt_uniform(&e->rng, e->obj_lo[1], e->obj_hi[1]);
        int ok = 1;
        for (int k = 0; k < b; k++) if (!(hypot(x - xy[2 * k], y - xy[2 * k + 1]) > 0.06)) ok = 0;
        if (!(hypot(x - e->tip_init[0], y - e->tip_init[1]) > 0.06)) ok = 0;
        if (ok) { xy[2 * b] = x; xy[2 * b + 1] = y; break; }
      }
    }
    place_blocks(e, xy);
    if (e->task == PMGO_BLOCK_REARRANGE) {
      /* kuka_multi_step_envs.py:174-189: one table target per block, clear of every block and earlier target */
      double txy[2 * MAXBLK];
      for (int b = 0; b < e->nb; b++) {
        for (;;) {
          double x = mt_uniform(&e->rng, e->tgt_lo[0], e->tgt_hi[0]);
          double y = mt_uniform(&e->rng, e->tgt_lo[1], e->tgt_hi[1]);
          int ok = 1;
          for (int k = 0; k < b; k++) if (!(hypot(x - txy[2 * k], y - txy[2 * k + 1]) > 0.06)) ok = 0;
          for (int k = 0; k < e->nb; k++) if (!(hypot(x - xy[2 * k], y - xy[2 * k + 1]) > 0.06)) ok = 0;
          if (ok) { txy[2 * b] = x; txy[2 * b + 1] = y; break; }
        }
        set3(e->goal + 3 * b, txy[2 * b], txy[2 * b + 1], 0.175);
        set3(e->last_targets[b], txy[2 * b], txy[2 * b + 1], 0.175);
      }
      if (e->cur) {
        /* kuka_multi_step_envs.py:197-212: level = np_random.choice(num_curriculum, p=curriculum_prob), then the
         * level + 1 blocks to move = sort(np_random.choice(arange(nb), size=level + 1, replace=False)), which in
         * numpy's legacy RandomState is permutation(nb)[:level + 1] (a shuffle of arange(nb)) */
        const int level = draw_curriculum_level(e);
        int64_t perm[MAXBLK];
        for (int k = 0; k < e->nb; k++) perm[k] = k;
        pmgo_rng_shuffle(e, perm, e->nb);
        int mask = 0;
        for (int k = 0; k <= level; k++) mask |= 1 << (int)perm[k];
        e->sub_goal_ind = mask;
        if (e->cur_update) { e->cur_count[level] += 1; update_curriculum_prob(e); }
        /* the goal words of the moved blocks hold their targets (what a spawn row carries, pmgo_reset_with) */
        int j = 0;
        for (int i = 0; i < e->nb; i++) if ((mask >> i) & 1) copy3(e->goal + 3 * i, e->last_targets[j++]);
      }
      write_obs(e, obs_out);
      return;
    }
    /* kuka_multi_step_envs.py:34-63 */
    int64_t order[MAXBLK];
    for (int k = 0; k < e->nb; k++) order[k] = k;
    pmgo_rng_shuffle(e, order, e->nb);
    double bx, by;
    for (;;) {
      bx = mt_uniform(&e->rng, e->tgt_lo[0], e->tgt_hi[0]);
      by = mt_uniform(&e->rng, e->tgt_lo[1], e->tgt_hi[1]);
      int ok = 1;
      for (int k = 0; k < e->nb; k++) if (!(hypot(bx - xy[2 * k], by - xy[2 * k + 1]) > 0.08)) ok = 0;
      if (ok) break;
    }
    for (int k = 0; k < e->nb; k++) {
      e->last_order[k] = (int)order[k];
      set3(e->last_targets[k], bx, by, k == 0 ? 0.175 : 0.175 + 0.03 * k);
      copy3(e->goal + 3 * e->last_order[k], e->last_targets[k]);
    }
    if (e->grip) { copy3(e->goal + 3 * e->nb, e->last_targets[e->nb - 1]); e->goal[3 * e->nb + 3] = 0.03; } /* :75-77 */
    if (e->cur) {
      /* kuka_multi_step_envs.py:127-134 */
      const int level = draw_curriculum_level(e);
      e->sub_goal_ind = e->grip ? 2 * level + 1 : level;
      if (e->cur_update) { e->cur_count[level] += 1; update_curriculum_prob(e); }
    }
  } else {
    /* kuka_single_step_base_env.py:104-148 */
    double center[3];
    copy3(center, e->tip_init);
    if (e->has_obj) {
      double x = e->tip_init[0], y = e->tip_init[1];
      while (hypot(x - e->tip_init[0], y - e->tip_init[1]) < 0.1) {
        x = mt_uniform(&e->rng, e->obj_lo[0], e->obj_hi[0]);
        y = mt_uniform(&e->rng, e->obj_lo[1], e->obj_hi[1]);
      }
      xy[0] = x; xy[1] = y;
      place_blocks(e, xy);
      set3(center, x, y, e->blk_spawn_z);
    }
    for (;;) {
      for (int k = 0; k < 3; k++) e->goal[k] = mt_uniform(&e->rng, e->tgt_lo[k], e->tgt_hi[k]);
      double d[3];
      sub3(d, e->goal, center);
      if (sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]) > 0.1) break;
    }
    if (!e->target_in_air) e->goal[2] = e->blk_spawn_z;
    else if (e->grasping) { if (mt_uniform(&e->rng, 0, 1) >= 0.5) e->goal[2] = e->blk_spawn_z; }
  }
  write_obs(e, obs_out);
}

void pmgo_reset_with(PmgoEnv* e, const double* spawn, double* obs_out) {
  robot_reset(e);
  e->elapsed = 0;
  /* spawn rows of a curriculum env carry the sub-goal index behind the goal */
  e->sub_goal_ind = e->cur ? (int)spawn[2 * e->nb + e->dims[3]] : -1;
  if (e->cur) e->cur_level = e->grip ? e->sub_goal_ind >> 1 : e->sub_goal_ind;
  place_blocks(e, spawn);
  memcpy(e->goal, spawn + 2 * e->nb, sizeof(double) * e->dims[3]);
  if (e->cur && e->task == PMGO_BLOCK_REARRANGE) {
    /* the index word is the mask of the blocks to move; their goal words hold the targets, in order */
    int j = 0;
    for (int i = 0; i < e->nb; i++) if ((e->sub_goal_ind >> i) & 1) copy3(e->last_targets[j++], e->goal + 3 * i);
    e->cur_level = j - 1;
  }
  if (e->task == PMGO_BLOCK_STACK) {
    /* recover order/targets from the goal: level k <=> z = 0.175 + 0.03 k */
    for (int b = 0; b < e->nb; b++) {
      int k = (int)floor((e->goal[3 * b + 2] - 0.175) / 0.03 + 0.5);
      e->last_order[k] = b;
      copy3(e->last_targets[k], e->goal + 3 * b);
    }
  }
  write_obs(e, obs_out);
}

void pmgo_compute_reward(const double* ag, const double* dg, int64_t n, int g, double thr, int binary,
                         double* reward, uint8_t* achieved) {
  for (int64_t i = 0; i < n; i++) {
    double s = 0;
    for (int k = 0; k < g; k++) { double d = ag[i * g + k] - dg[i * g + k]; s += d * d; }
    double d = sqrt(s);
    int na = d > thr;
    reward[i] = binary ? -(double)(float)na : -d;
    achieved[i] = (uint8_t)!na;
  }
}

void pmgo_step(PmgoEnv* e, const double* a, double* obs_out, double* reward, int* done, int* goal_achieved) {
  /* kuka.py:167-225 */
  if (e->grasping) {
    double grip = (a[e->adim - 1] + 1.0) * (GRIPPER_ABS_LIMIT / 2);
    for (int d = 7; d < ND; d++) { e->mot_target[d] = grip; e->mot_maximp[d] = FINGER_FORCE * OUTER_DT; }
  }
  if (e->jc) {
    /* kuka.py:204-206: joint_state_target += 0.05 a[:7], no clipping, no IK */
    for (int d = 0; d < 7; d++) { e->mot_target[d] += a[d] * 0.05; e->mot_maximp[d] = ARM_FORCE * OUTER_DT; }
  } else {
    for (int k = 0; k < 3; k++) {
      e->ee_target[k] += a[k] * 0.01;
      if (e->ee_target[k] < EE_LOWER[k]) e->ee_target[k] = EE_LOWER[k];
      if (e->ee_target[k] > EE_UPPER[k]) e->ee_target[k] = EE_UPPER[k];
    }
    double qik[ND];
    pmgo_ik(e->q, e->ee_target, EE_FIXED_QUAT, 40, 1e-5, qik);
    for (int d = 0; d < 7; d++) { e->mot_target[d] = qik[d]; e->mot_maximp[d] = ARM_FORCE * OUTER_DT; }
  }
  for (int c = 0; c < CALLS_PER_ENV_STEP; c++) pmgo_step_simulation(e);
  write_obs(e, obs_out);
  uint8_t ok;
  pmgo_compute_reward(obs_out + e->dims[0] + e->dims[1], obs_out + e->dims[0] + e->dims[1] + e->dims[2], 1, e->dims[2],
                      e->thr, e->binary, reward, &ok);
  *goal_achieved = ok;
  e->elapsed++;
  *done = e->elapsed >= e->max_steps; /* gym TimeLimit */
}

/* ------------------------------------------------------------------------------------------ */
/* state access + diagnostics                                                                 */
/* ------------------------------------------------------------------------------------------ */
int pmgo_state_size(const PmgoEnv* e) { return 9 + 9 + 3 + 7 + 9 + 9 + 13 * e->nb + e->dims[3] + ((e->td || e->cur) ? 1 : 0) + 1; }
void pmgo_get_state(const PmgoEnv* e, double* o) {
  memcpy(o, e->q, 72); o += 9; memcpy(o, e->qd, 72); o += 9;
  memcpy(o, e->ee_target, 24); o += 3; memcpy(o, e->rest_pose, 56); o += 7;
  memcpy(o, e->mot_target, 72); o += 9; memcpy(o, e->mot_maximp, 72); o += 9;
  for (int b = 0; b < e->nb; b++) {
    copy3(o, e->bpos[b]); memcpy(o + 3, e->bquat[b], 32); copy3(o + 7, e->bv[b]); copy3(o + 10, e->bw[b]); o += 13;
  }
  memcpy(o, e->goal, sizeof(double) * e->dims[3]); o += e->dims[3];
  if (e->td || e->cur) *o++ = e->sub_goal_ind;
  *o = e->elapsed;
}
void pmgo_set_state(PmgoEnv* e, const double* o) {
  memcpy(e->q, o, 72); o += 9; memcpy(e->qd, o, 72); o += 9;
  memcpy(e->ee_target, o, 24); o += 3; memcpy(e->rest_pose, o, 56); o += 7;
  memcpy(e->mot_target, o, 72); o += 9; memcpy(e->mot_maximp, o, 72); o += 9;
  for (int b = 0; b < e->nb; b++) {
    copy3(e->bpos[b], o); memcpy(e->bquat[b], o + 3, 32); copy3(e->bv[b], o + 7); copy3(e->bw[b], o + 10); o += 13;
  }
  memcpy(e->goal, o, sizeof(double) * e->dims[3]); o += e->dims[3];
  if (e->td || e->cur) e->sub_goal_ind = (int)*o++;
  e->elapsed = (int)*o;
  if (e->task == PMGO_BLOCK_STACK)
    for (int b = 0; b < e->nb; b++) {
      int k = (int)floor((e->goal[3 * b + 2] - 0.175) / 0.03 + 0.5);
      e->last_order[k] = b;
      copy3(e->last_targets[k], e->goal + 3 * b);
    }
  if (e->cur && e->task == PMGO_BLOCK_REARRANGE) {
    int j = 0;
    for (int i = 0; i < e->nb; i++) if ((e->sub_goal_ind >> i) & 1) copy3(e->last_targets[j++], e->goal + 3 * i);
  }
  memset(e->man, 0, sizeof e->man);
}

void pmgo_poke_state(PmgoEnv* e, const double* o) { /* like set_state, but the contact caches survive */
  Manifold keep[MAX_PAIRS];
  memcpy(keep, e->man, sizeof keep);
  pmgo_set_state(e, o);
  memcpy(e->man, keep, sizeof keep);
}

void pmgo_mass_matrix_inverse(PmgoEnv* e, double* out) {
  double qdd[ND], tau[ND], col[ND];
  memset(tau, 0, sizeof tau);
  aba(e, tau, qdd);
  for (int d = 0; d < ND; d++) {
    memset(tau, 0, sizeof tau);
    tau[d] = 1;
    impulse_response(e, NULL, tau, col);
    for (int r = 0; r < ND; r++) out[r * ND + d] = col[r];
  }
}

int pmgo_get_contacts(const PmgoEnv* e, double* out, int max) {
  int n = 0;
  for (int k = 0; k < e->npair; k++) {
    const double *pa, *Ra, *pb, *Rb;
    geom_pose(e, &e->pairs[k].a, &pa, &Ra);
    geom_pose(e, &e->pairs[k].b, &pb, &Rb);
    for (int i = 0; i < e->man[k].n && n < max; i++, n++) {
      double* o = out + 11 * n;
      o[0] = k;
      matvec3(Ra, e->man[k].lA[i], o + 1); add3(o + 1, o + 1, pa);
      matvec3(Rb, e->man[k].lB[i], o + 4); add3(o + 4, o + 4, pb);
      copy3(o + 7, e->man[k].nB[i]);
      o[10] = e->man[k].dist[i];
    }
  }
  return n;
}

/* ------------------------------------------------------------------------------------------ */
/* multi-threaded rollout driver for the CPU baseline                                         */
/* ------------------------------------------------------------------------------------------ */
typedef struct { PmgoEnv** envs; int n_env, lo, hi, n_steps; const double* actions; } Job;
static void* job_main(void* arg) {
  Job* j = (Job*)arg;
  double obs[PMGO_MAX_OBS + 64], r; int done, ok;
  for (int t = 0; t < j->n_steps; t++)
    for (int i = j->lo; i < j->hi; i++) {
      int ad = j->envs[i]->adim;
      pmgo_step(j->envs[i], j->actions + ((size_t)t * j->n_env + i) * ad, obs, &r, &done, &ok);
      if (done) pmgo_reset(j->envs[i], obs);
    }
  return NULL;
}
double pmgo_bench_rollout(PmgoEnv** envs, int n_env, const double* actions, int n_steps, int n_threads) {
  if (n_threads < 1) n_threads = 1;
  if (n_threads > n_env) n_threads = n_env;
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * n_threads);
  Job* jobs = (Job*)malloc(sizeof(Job) * n_threads);
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int k = 0; k < n_threads; k++) {
    jobs[k].envs = envs; jobs[k].n_env = n_env; jobs[k].n_steps = n_steps; jobs[k].actions = actions;
    jobs[k].lo = (int)((long)n_env * k / n_threads); jobs[k].hi = (int)((long)n_env * (k + 1) / n_threads);
    pthread_create(&th[k], NULL, job_main, &jobs[k]);
  }
  for (int k = 0; k < n_threads; k++) pthread_join(th[k], NULL);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  free(th); free(jobs);
  return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}
