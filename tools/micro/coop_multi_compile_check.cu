// Compile check (not part of libpmg.so): instantiates the multi-block cooperative step for sm_100a so that
// ptxas reports its register / stack use.  nvcc -gencode arch=compute_100a,code=sm_100a -Xptxas -v -c this file.
#include "../../pybullet_multigoal_gym_b200/csrc/pmg_coop.cuh"

using namespace pmg;

template <int NBLK>
__global__ void __launch_bounds__(32, 4) step_kernel_coop_multi(StepIO io) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane32 = threadIdx.x & 31, grp = lane32 >> 3;
  float* lane_consts = reinterpret_cast<float*>(smem);
  if (lane32 < coop::GL) coop::fill_lane_constants(lane_consts + lane32 * coop::LC_W, lane32);
  __syncwarp();
  const int env = blockIdx.x * (32 / coop::GL) + grp;
  if (env >= io.batch) return;
  coop::Grp g;
  g.lane = lane32 & (coop::GL - 1); g.shift = grp * coop::GL; g.mask = 0xffu << g.shift;
  coop::EnvSmemT<NBLK>& sm = reinterpret_cast<coop::EnvSmemT<NBLK>*>(smem + 1024)[grp];
  coop::step_env_multi<NBLK>(g, sm, lane_consts, io, env);
}
template __global__ void step_kernel_coop_multi<4>(StepIO);
template __global__ void step_kernel_coop_multi<2>(StepIO);
