// Instruction-cache cliff probe: a loop whose body is N straight-line FFMAs (8 independent chains),
// launched like the step kernel (one warp per block, ~2 warps per SM).  Prints cycles per instruction.
#include <cstdio>
#include <cuda_runtime.h>
template <int N>
__global__ void probe(float* out, int iters, long long* cyc) {
  float r[8];
  for (int k = 0; k < 8; k++) r[k] = threadIdx.x * 0.001f + k;
  float a = out[0], b = out[1];
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < N; i++) r[i & 7] = fmaf(r[i & 7], a, b + (float)(i & 1));
  }
  long long t1 = clock64();
  float s = 0;
  for (int k = 0; k < 8; k++) s += r[k];
  out[2 + blockIdx.x * 32 + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int N> void run(float* d, long long* dc, int blocks) {
  int iters = 400000 / N + 8;
  probe<N><<<blocks, 32>>>(d, iters, dc);
  cudaDeviceSynchronize();
  probe<N><<<blocks, 32>>>(d, iters, dc);
  cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
  printf("blocks %4d body %6d instr (%4d KB): %.3f cycles/instr\n", blocks, N, N * 16 / 1024, (double)c / ((double)N * iters));
}
int main() {
  float* d; long long* dc;
  cudaMalloc(&d, 1 << 22); cudaMemset(d, 0, 1 << 22); cudaMalloc(&dc, 8);
  for (int blocks : {256, 2048}) {
    run<512>(d, dc, blocks); run<1024>(d, dc, blocks); run<2048>(d, dc, blocks); run<3072>(d, dc, blocks); run<4096>(d, dc, blocks);
    run<6144>(d, dc, blocks); run<8192>(d, dc, blocks); run<12288>(d, dc, blocks); run<16384>(d, dc, blocks);
  }
  return 0;
}
