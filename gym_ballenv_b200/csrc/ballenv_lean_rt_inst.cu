// Instantiations of the thread-per-environment kernel with run-time obstacle counts (ballenv_lean_rt.cuh), one
// translation unit per (window, lanes per environment):
//   -DBALLENV_W=5|10 -DBALLENV_KS=-1 -DBALLENV_KD=-1 -DBALLENV_G=1|2 -DBALLENV_NAME=launch_lean_w.._rt_g..
//   -DBALLENV_SMEM_NAME=lean_rt_smem_w.._g..   (bytes of dynamic shared memory a block needs for given counts)
#include <cuda_runtime.h>
#include <stdlib.h>

#include "ballenv_lean_rt.cuh"

#if !defined(BALLENV_W) || !defined(BALLENV_KS) || !defined(BALLENV_KD) || !defined(BALLENV_G) || !defined(BALLENV_NAME)
#error "compile with -DBALLENV_W -DBALLENV_KS -DBALLENV_KD -DBALLENV_G -DBALLENV_NAME"
#endif

namespace ballenv {

template <bool kRollout, bool kPolicy = false>
static void launch_lean(const Params& p, unsigned grid, cudaStream_t s) {
  auto kern = ballenv_lean_kernel<BALLENV_W, BALLENV_KS, BALLENV_KD, BALLENV_G, kRollout, kPolicy>;
  static bool configured[64] = {};   // per device: function attributes belong to the device's copy of the kernel
  int dev = 0;
  cudaGetDevice(&dev);
  constexpr int kWarps = kLeanEnvsPerBlock * BALLENV_G / 32;
  // dynamic shared memory: (policy in the loop) the block's copy of the weights; (run-time obstacle counts) the warps'
  // regions, sized by the configuration
#if BALLENV_KS < 0
  const size_t dyn = (size_t)LeanMem<BALLENV_W, BALLENV_KS, BALLENV_KD, BALLENV_G>::bytes(p.cfg.ks, p.cfg.kd) * kWarps +
                     (kPolicy ? 4 * lean::policy_smem_floats(4 + BALLENV_W * BALLENV_W, p.pol_hidden) : 0);
  const int dyn_max = 220 * 1024;
#else
  const size_t dyn = kPolicy ? 4 * lean::policy_smem_floats(4 + BALLENV_W * BALLENV_W, p.pol_hidden) : 0;
  const int dyn_max = kPolicy ? 192 * 1024 - (int)sizeof(LeanWarp<BALLENV_W, BALLENV_KS, BALLENV_KD, BALLENV_G>) * kWarps : 0;
#endif
  if (dev >= 0 && dev < 64 && !configured[dev]) {
    // seven blocks of 64 environments with their obstacle slices in shared memory: ask for the whole array
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (dyn_max > 0) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn_max);
    configured[dev] = true;
  }
  // programmatic stream serialisation: the grid may begin while its predecessor on the stream drains; the kernel
  // itself waits (griddepcontrol.wait) before it reads anything (BALLENV_NO_PDL=1 launches it the plain way)
  static const bool no_pdl = getenv("BALLENV_NO_PDL") != nullptr && getenv("BALLENV_NO_PDL")[0] == '1';
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3(grid);
  lc.blockDim = dim3(kLeanEnvsPerBlock * BALLENV_G);
  lc.dynamicSmemBytes = dyn;
  lc.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  lc.attrs = attr;
  lc.numAttrs = no_pdl ? 0 : 1;
#ifdef LEAN_TRACE
  // profiling build (tools/lean_trace.py): the stamp buffer's address comes from the environment, the launch number
  // rides in Params::debug
  static unsigned long long counter = 0;
  if (counter == 0) {
    unsigned long long* buf = (unsigned long long*)strtoull(getenv("BALLENV_TRACE_PTR") ? getenv("BALLENV_TRACE_PTR") : "0", nullptr, 0);
    cudaMemcpyToSymbol(lean::lean_trace_buf, &buf, sizeof(buf));
  }
  Params q = p;
  q.debug = (int)(counter++ & 63);
  cudaLaunchKernelEx(&lc, kern, q);
  return;
#endif
  cudaLaunchKernelEx(&lc, kern, p);
}

// grid = blocks of kLeanEnvsPerBlock environments
void BALLENV_NAME(const Params& p, unsigned grid, cudaStream_t s) {
  if (p.n_steps > 1) launch_lean<true>(p, grid, s);
  else launch_lean<false>(p, grid, s);
}

// dynamic shared memory of one block for these obstacle counts (the dispatcher asks before it picks this kernel)
size_t BALLENV_SMEM_NAME(int ks, int kd) {
  return (size_t)LeanMem<BALLENV_W, BALLENV_KS, BALLENV_KD, BALLENV_G>::bytes(ks, kd) * (kLeanEnvsPerBlock * BALLENV_G / 32);
}

#ifdef BALLENV_POLICY_NAME
// the rollout loop with the policy inside (ballenv_rollout_policy): any number of steps
void BALLENV_POLICY_NAME(const Params& p, unsigned grid, cudaStream_t s) { launch_lean<true, true>(p, grid, s); }
#endif

}  // namespace ballenv
