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
template <int kLW>
static void launch_rollout(const Params& p, unsigned grid, cudaStream_t s) {
  auto kern = ballenv_kernel<BALLENV_T, BALLENV_W, true, true, kLW>;
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
#endif

void BALLENV_NAME(const Params& p, unsigned grid, cudaStream_t s) {
#if BALLENV_FAST
  // at most six quads per environment (the reference's default 13 + 5 obstacles): blocks of 6 obstacle warps, five
  // of them resident per SM instead of four
  const bool six = p.n_slot <= 6 * kEnvsPerBlock;
  if (p.n_steps > 1) {
    if (six) launch_rollout<6>(p, grid, s);
    else launch_rollout<kLanes>(p, grid, s);
  } else if (six) {
    ballenv_kernel<BALLENV_T, BALLENV_W, true, false, 6><<<grid, 32 + 32 * 6, 0, s>>>(p);
  } else {
    ballenv_kernel<BALLENV_T, BALLENV_W, true, false><<<grid, kBlock, 0, s>>>(p);
  }
#else
  ballenv_kernel<BALLENV_T, BALLENV_W, false, false><<<grid, kBlock, 0, s>>>(p);
#endif
}
}  // namespace ballenv
