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
void BALLENV_NAME(const Params& p, unsigned grid, cudaStream_t s) {
#if BALLENV_FAST
  if (p.n_steps > 1) {
    // The rollout kernel stages its rows in shared memory (38 KB per block for WINDOW = 10): ask for the smallest
    // carve-out that keeps kMinBlocks blocks resident, and no more - the rest of the array is the L1 the spilled
    // loop state lives in.
    static int carveout = -1;
    if (carveout < 0) {
      cudaFuncAttributes fa;
      cudaFuncGetAttributes(&fa, ballenv_kernel<BALLENV_T, BALLENV_W, true, true>);
      const size_t need = (size_t)kMinBlocks<BALLENV_T> * (fa.sharedSizeBytes + 1024);
      carveout = (int)((need * 100 + 228 * 1024 - 1) / (228 * 1024));
      if (carveout > 100) carveout = 100;
    }
    cudaFuncSetAttribute(ballenv_kernel<BALLENV_T, BALLENV_W, true, true>, cudaFuncAttributePreferredSharedMemoryCarveout,
                         carveout);
    ballenv_kernel<BALLENV_T, BALLENV_W, true, true><<<grid, kBlock, 0, s>>>(p);
  } else ballenv_kernel<BALLENV_T, BALLENV_W, true, false><<<grid, kBlock, 0, s>>>(p);
#else
  ballenv_kernel<BALLENV_T, BALLENV_W, false, false><<<grid, kBlock, 0, s>>>(p);
#endif
}
}  // namespace ballenv
