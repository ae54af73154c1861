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
  if (p.n_steps > 1) ballenv_kernel<BALLENV_T, BALLENV_W, true, true><<<grid, kBlock, 0, s>>>(p);
  else ballenv_kernel<BALLENV_T, BALLENV_W, true, false><<<grid, kBlock, 0, s>>>(p);
#else
  ballenv_kernel<BALLENV_T, BALLENV_W, false, false><<<grid, kBlock, 0, s>>>(p);
#endif
}
}  // namespace ballenv
