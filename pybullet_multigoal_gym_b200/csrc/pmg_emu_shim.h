// pmg_emu_shim.h -- lets a host C++ compiler (g++) build the device code of csrc/ unchanged.
//
// TEST INFRASTRUCTURE: only tests/emu/ defines PMG_EMULATE.  The lane-cooperative kernel
// (pmg_coop.cuh) is written against a tiny group interface (shuffles, ballot, sync over the 8 lanes of
// one environment); on the GPU those are __shfl_sync / __syncwarp, here they are provided by a
// coroutine scheduler that runs the 8 lanes of one environment in lockstep (tests/emu/pmg_coop_emu.cpp).
// This is how the cooperative algorithm is debugged on a machine without a GPU; it is never a product
// path (the Python package only ever loads libpmg.so and fails without it).
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__
#define __constant__ static const
#define __align__(n) alignas(n)
#define __launch_bounds__(...)

struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) float2 { float x, y; };
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline int __ffs(unsigned v) { return v ? __builtin_ctz(v) + 1 : 0; }
static inline int __clz(unsigned v) { return v ? __builtin_clz(v) : 32; }
static inline int atomicAdd(int* p, int v) { int o = *p; *p += v; return o; }
static inline void __threadfence_system() {}
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }

// group primitives implemented by the emulator
namespace pmg_emu {
int lane();
float shfl(float v, int src);
unsigned ballot(bool pred);
void sync();
}  // namespace pmg_emu
