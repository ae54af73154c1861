// One explicit instantiation of the fused kernel per translation unit (compiled once per
// -DBALLENV_T / -DBALLENV_W pair so the instantiations build in parallel).
#include <cuda_runtime.h>

#include "ballenv_kernels.cuh"

#ifndef BALLENV_T
#error "compile with -DBALLENV_T=float|double -DBALLENV_W=0|5|10 -DBALLENV_FAST=0|1 -DBALLENV_NAME=..."
#endif
#ifndef BALLENV_FAST
#define BALLENV_FAST 0
#endif

namespace ballenv {

#if BALLENV_FAST
// The rollout kernel stages its rows in shared memory (13 KB for WINDOW = 10): ask for the smallest carve-out that keeps
// its resident blocks, and no more - the rest of the array is the L1 the spilled loop state lives in.
template <int kLW, int kSQ>
static void launch_rollout(const Params& p, unsigned grid, cudaStream_t s) {
  auto kern = ballenv_kernel<BALLENV_T, BALLENV_W, true, true, kLW, kSQ>;
  static int carveout = -1;
  if (carveout < 0) {
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, kern);
    const size_t need = (size_t)kMinBlocks<BALLENV_T, kLW> * (fa.sharedSizeBytes + 1024);
    carveout = (int)((need * 100 + 228 * 1024 - 1) / (228 * 1024));
    if (carveout > 100) carveout = 100;
  }
  cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, carveout);
  kern<<<grid, 32 + 32 * kLW, 0, s>>>(p);
}
template <int kLW, int kSQ>
static void launch_fast(const Params& p, unsigned grid, cudaStream_t s) {
  if (p.n_steps > 1) launch_rollout<kLW, kSQ>(p, grid, s);
  else ballenv_kernel<BALLENV_T, BALLENV_W, true, false, kLW, kSQ><<<grid, 32 + 32 * kLW, 0, s>>>(p);
}
#endif

void BALLENV_NAME(const Params& p, unsigned grid, cudaStream_t s) {
#if BALLENV_FAST
  // The smallest block that holds the configuration's obstacle threads: more resident blocks hide more of a step's
  // phase chain.  6 obstacle warps for at most six quads per environment, else 8; p.sq = 2 (ballenv_capi.cu: launch,
  // single-step launches of small configurations only): a static-quad thread takes two quads, 4 obstacle warps.
  const int warps = (p.n_slot + 31) / 32;
  if (p.sq == 2 && p.n_steps == 1 && warps <= 4) {
    ballenv_kernel<BALLENV_T, BALLENV_W, true, false, 4, 2><<<grid, 32 + 32 * 4, 0, s>>>(p);
  } else if (warps <= 6) {
    launch_fast<6, 1>(p, grid, s);
  } else {
    launch_fast<kLanes, 1>(p, grid, s);
  }
#else
  ballenv_kernel<BALLENV_T, BALLENV_W, false, false><<<grid, kBlock, 0, s>>>(p);
#endif
}
}  // namespace ballenv
