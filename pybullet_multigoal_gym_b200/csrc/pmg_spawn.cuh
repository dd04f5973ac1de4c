// pmg_spawn.cuh -- device-side reset sampling: object / goal poses drawn inside the reset kernel.
//
// The reference samples every reset from gym's np_random (MT19937) with data-dependent draw counts
// (kuka_single_step_base_env.py:104-148, kuka_multi_step_base_env.py:223-240, kuka_multi_step_envs.py:34-87,
// :174-189); pmg_reset reproduces that stream bit-exactly on the host.  This is the throughput path
// ("spawn NULL => device Philox"): the SAME sampling rules -- boxes, rejection tests, order
// shuffle, z rules -- drawn from a counter-based Philox4x32-10 stream per (seed, global env index, episode), in
// float32, so an environment resets itself on the device with no host round trip (pmg_reset_device, auto-reset).
// It is a different random stream from the reference's, by construction; oracle/device_rng_oracle.py restates it
// in numpy and the tests compare the spawn rows bit-exactly (every multiply / add below is individually rounded,
// no fused multiply-add, so that numpy float32 arithmetic reproduces it).
#pragma once

#include <stdint.h>

#ifndef PMG_EMULATE
#define PMG_HD __device__ __forceinline__
#define PMG_FMUL(a, b) __fmul_rn((a), (b))
#define PMG_FADD(a, b) __fadd_rn((a), (b))
#else
#define PMG_HD static inline
#define PMG_FMUL(a, b) ((float)((a) * (b)))
#define PMG_FADD(a, b) ((float)((a) + (b)))
#endif

namespace pmg {
namespace spawn {

// Philox4x32-10 (Salmon et al., SC'11), the published round constants.
struct Philox {
  uint32_t key[2], ctr[4], out[4];
  int have;  // unread words in out
};

PMG_HD void philox_block(const uint32_t ctr[4], const uint32_t key_in[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key_in[0], k1 = key_in[1];
  for (int r = 0; r < 10; r++) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// stream of one (seed, env, episode): counter = (block number, episode, env lo, env hi), key = seed
PMG_HD void stream_init(Philox& s, uint64_t seed, int64_t env, uint32_t episode) {
  s.key[0] = (uint32_t)seed; s.key[1] = (uint32_t)(seed >> 32);
  s.ctr[0] = 0; s.ctr[1] = episode; s.ctr[2] = (uint32_t)(uint64_t)env; s.ctr[3] = (uint32_t)((uint64_t)env >> 32);
  s.have = 0;
}
PMG_HD uint32_t next_u32(Philox& s) {
  if (s.have == 0) { philox_block(s.ctr, s.key, s.out); s.ctr[0]++; s.have = 4; }
  const uint32_t v = s.out[4 - s.have];
  s.have--;
  return v;
}
PMG_HD float uniform(Philox& s, float lo, float hi) {  // lo + (hi - lo) * u, u = 24 random bits in [0, 1)
  const float u = (float)(next_u32(s) >> 8) * (1.0f / 16777216.0f);
  return PMG_FADD(lo, PMG_FMUL(PMG_FADD(hi, -lo), u));
}
PMG_HD float dist2(float ax, float ay, float bx, float by) {
  const float dx = PMG_FADD(ax, -bx), dy = PMG_FADD(ay, -by);
  return PMG_FADD(PMG_FMUL(dx, dx), PMG_FMUL(dy, dy));
}

struct Bounds { float tip[3], obj_lo[2], obj_hi[2], tgt_lo[3], tgt_hi[3]; };  // kuka.py:35-51 per task (pmg_create)

// Sampling boxes of a task: kuka.py:35-51 with obj_range = target_range = 0.15 (kuka_single_step_envs.py,
// kuka_multi_step_envs.py:29), tip start 1 mm above the table for Push / BlockRearrange (kuka.py:37-38).  Host only.
// Slide (task 5, kuka_single_step_envs.py:49-59, kuka_single_step_base_env.py:66-69): obj_range 0.1, target_range 0.2,
// targets 0.4 further along -x (beyond the arm's reach).
inline void task_bounds(int task, double tip[3], double obj_lo[3], double obj_hi[3], double tgt_lo[3], double tgt_hi[3]) {
  tip[0] = -0.52; tip[1] = 0.0; tip[2] = (task == 1 || task == 4 || task == 5) ? 0.175 + 0.001 : 0.25;
  const double obj_range = task == 5 ? 0.1 : 0.15, target_range = task == 5 ? 0.2 : 0.15;
  for (int k = 0; k < 3; k++) {
    obj_lo[k] = tip[k] - obj_range; obj_hi[k] = tip[k] + obj_range;
    tgt_lo[k] = tip[k] - target_range; tgt_hi[k] = tip[k] + target_range;
  }
  obj_lo[0] += 0.03; obj_hi[0] -= 0.03;
  tgt_lo[0] += 0.03; tgt_lo[2] = 0.175; tgt_hi[0] -= 0.03;
  if (task == 5) { tgt_lo[0] -= 0.4; tgt_hi[0] -= 0.4; }
}
inline Bounds to_bounds(const double tip[3], const double obj_lo[3], const double obj_hi[3], const double tgt_lo[3], const double tgt_hi[3]) {
  Bounds b;
  for (int k = 0; k < 3; k++) { b.tip[k] = (float)tip[k]; b.tgt_lo[k] = (float)tgt_lo[k]; b.tgt_hi[k] = (float)tgt_hi[k]; }
  for (int k = 0; k < 2; k++) { b.obj_lo[k] = (float)obj_lo[k]; b.obj_hi[k] = (float)obj_hi[k]; }
  return b;
}

constexpr int MAX_TRIES = 64;  // rejection loops are bounded on the device; the last candidate is kept (p < 1e-30)

// One reset's spawn row [block xy (2 nb) | goal (G)], the layout pmg_reset takes.  task: pmg_task ids.
// grip: grip-informed goal (block_stack).  Curriculum resets are host-only (their schedule lives on the host).
PMG_HD void sample_row(Philox& r, int task, int nb, int grip, const Bounds& b, float* out) {
  const float R01 = 0.01f, R006 = 0.0036f, R008 = 0.0064f;  // 0.1^2, 0.06^2, 0.08^2 as float32 literals
  const float Z0 = task == 5 ? 0.17f : 0.175f;  // spawn height of the object: cube 0.175, Slide's puck 0.170 (kuka_single_step_base_env.py:50,56)
  if (task == 3 || task == 4) {  // block_stack / block_rearrange
    for (int k = 0; k < nb; k++) {  // kuka_multi_step_base_env.py:223-240
      float x = 0.0f, y = 0.0f;
      for (int tries = 0; tries < MAX_TRIES; tries++) {
        x = uniform(r, b.obj_lo[0], b.obj_hi[0]); y = uniform(r, b.obj_lo[1], b.obj_hi[1]);
        bool ok = dist2(x, y, b.tip[0], b.tip[1]) > R006;
        for (int j = 0; j < k; j++) ok = ok && dist2(x, y, out[2 * j], out[2 * j + 1]) > R006;
        if (ok) break;
      }
      out[2 * k] = x; out[2 * k + 1] = y;
    }
    float* goal = out + 2 * nb;
    if (task == 4) {  // kuka_multi_step_envs.py:174-189: one table target per block
      for (int k = 0; k < nb; k++) {
        float x = 0.0f, y = 0.0f;
        for (int tries = 0; tries < MAX_TRIES; tries++) {
          x = uniform(r, b.tgt_lo[0], b.tgt_hi[0]); y = uniform(r, b.tgt_lo[1], b.tgt_hi[1]);
          bool ok = true;
          for (int j = 0; j < k; j++) ok = ok && dist2(x, y, goal[3 * j], goal[3 * j + 1]) > R006;
          for (int j = 0; j < nb; j++) ok = ok && dist2(x, y, out[2 * j], out[2 * j + 1]) > R006;
          if (ok) break;
        }
        goal[3 * k] = x; goal[3 * k + 1] = y; goal[3 * k + 2] = Z0;
      }
      return;
    }
    int order[5];  // kuka_multi_step_envs.py:37-40: a random stacking order (Fisher-Yates, as numpy's shuffle walks it)
    for (int k = 0; k < nb; k++) order[k] = k;
    for (int k = nb - 1; k > 0; k--) {
      const int j = (int)(((uint64_t)next_u32(r) * (uint64_t)(k + 1)) >> 32);
      const int t = order[k]; order[k] = order[j]; order[j] = t;
    }
    float bx = 0.0f, by = 0.0f;
    for (int tries = 0; tries < MAX_TRIES; tries++) {  // :45-53: the stack's base, clear of every block
      bx = uniform(r, b.tgt_lo[0], b.tgt_hi[0]); by = uniform(r, b.tgt_lo[1], b.tgt_hi[1]);
      bool ok = true;
      for (int j = 0; j < nb; j++) ok = ok && dist2(bx, by, out[2 * j], out[2 * j + 1]) > R008;
      if (ok) break;
    }
    for (int k = 0; k < nb; k++) {
      goal[3 * order[k]] = bx; goal[3 * order[k] + 1] = by;
      goal[3 * order[k] + 2] = PMG_FADD(Z0, PMG_FMUL(0.03f, (float)k));
    }
    if (grip) {  // :75-77
      goal[3 * nb] = bx; goal[3 * nb + 1] = by; goal[3 * nb + 2] = PMG_FADD(Z0, PMG_FMUL(0.03f, (float)(nb - 1))); goal[3 * nb + 3] = 0.03f;
    }
    return;
  }
  float cx = b.tip[0], cy = b.tip[1], cz = b.tip[2];  // kuka_single_step_base_env.py:104-148
  if (nb) {
    float x = 0.0f, y = 0.0f;
    for (int tries = 0; tries < MAX_TRIES; tries++) {
      x = uniform(r, b.obj_lo[0], b.obj_hi[0]); y = uniform(r, b.obj_lo[1], b.obj_hi[1]);
      if (!(dist2(x, y, b.tip[0], b.tip[1]) < R01)) break;
    }
    out[0] = x; out[1] = y;
    cx = x; cy = y; cz = Z0;
  }
  float g0 = 0.0f, g1 = 0.0f, g2 = 0.0f;
  for (int tries = 0; tries < MAX_TRIES; tries++) {
    g0 = uniform(r, b.tgt_lo[0], b.tgt_hi[0]); g1 = uniform(r, b.tgt_lo[1], b.tgt_hi[1]); g2 = uniform(r, b.tgt_lo[2], b.tgt_hi[2]);
    const float dz = PMG_FADD(g2, -cz);
    if (PMG_FADD(dist2(g0, g1, cx, cy), PMG_FMUL(dz, dz)) > R01) break;
  }
  if (task == 1 || task == 5) g2 = Z0;                      // push / slide: on the table (:138-139)
  else if (task == 2) { if (uniform(r, 0.0f, 1.0f) >= 0.5f) g2 = Z0; }  // pick_and_place: half of the goals on the table (:140-143)
  out[2 * nb] = g0; out[2 * nb + 1] = g1; out[2 * nb + 2] = g2;
}

}  // namespace spawn
}  // namespace pmg
