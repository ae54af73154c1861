// C ABI (include/ballenv.h) over the sm_100a kernels in ballenv_kernels.cuh.
// Host-side runtime: config validation, SoA arena layout, draw tapes, launches, host-buffer staging.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>

#include "ballenv_kernels.cuh"
#include "ballenv_features.cuh"
#include "ballenv_lean.cuh"
#include "ballenv_patches.cuh"
#include "ballenv_a2c.cuh"
#include "ballenv_reset_fixed.cuh"

using namespace ballenv;

namespace ballenv {  // ballenv_inst.cu, one translation unit per instantiation
void launch_f32_w5(const Params&, unsigned, cudaStream_t);
void launch_f32_w10(const Params&, unsigned, cudaStream_t);
void launch_f32_wany(const Params&, unsigned, cudaStream_t);
void launch_f64_w5(const Params&, unsigned, cudaStream_t);
void launch_f64_w10(const Params&, unsigned, cudaStream_t);
void launch_f64_wany(const Params&, unsigned, cudaStream_t);
void launch_f32_w5_fast(const Params&, unsigned, cudaStream_t);
void launch_f32_w10_fast(const Params&, unsigned, cudaStream_t);
void launch_f32_wany_fast(const Params&, unsigned, cudaStream_t);
// ballenv_lean_inst.cu: thread-per-environment kernels, one per (window, static, dynamic obstacle count)
// and lanes per environment (g1 / g2)
#define BALLENV_LEAN_DECL(w, ks, kd)                                          \
  void launch_lean_w##w##_s##ks##_d##kd##_g1(const Params&, unsigned, cudaStream_t); \
  void launch_lean_w##w##_s##ks##_d##kd##_g2(const Params&, unsigned, cudaStream_t);
// the same with the policy inside the rollout loop (ballenv_rollout_policy)
#define BALLENV_LEAN_POLICY_DECL(w, ks, kd)                                          \
  void launch_lean_policy_w##w##_s##ks##_d##kd##_g1(const Params&, unsigned, cudaStream_t); \
  void launch_lean_policy_w##w##_s##ks##_d##kd##_g2(const Params&, unsigned, cudaStream_t);
BALLENV_LEAN_DECL(5, 13, 5)
BALLENV_LEAN_DECL(10, 13, 5)
BALLENV_LEAN_DECL(10, 8, 24)
BALLENV_LEAN_DECL(5, 8, 24)
// the same with the obstacle counts read at run time (any counts whose lists fit)
void launch_lean_w5_rt_g1(const Params&, unsigned, cudaStream_t);
void launch_lean_w5_rt_g2(const Params&, unsigned, cudaStream_t);
void launch_lean_w10_rt_g1(const Params&, unsigned, cudaStream_t);
void launch_lean_w10_rt_g2(const Params&, unsigned, cudaStream_t);
void launch_lean_policy_w5_rt_g1(const Params&, unsigned, cudaStream_t);
void launch_lean_policy_w5_rt_g2(const Params&, unsigned, cudaStream_t);
void launch_lean_policy_w10_rt_g1(const Params&, unsigned, cudaStream_t);
void launch_lean_policy_w10_rt_g2(const Params&, unsigned, cudaStream_t);
size_t lean_rt_smem_w5_g1(int ks, int kd);
size_t lean_rt_smem_w5_g2(int ks, int kd);
size_t lean_rt_smem_w10_g1(int ks, int kd);
size_t lean_rt_smem_w10_g2(int ks, int kd);
constexpr int kLeanRtMaxObstacles = 64;   // ballenv_lean_rt.cuh
BALLENV_LEAN_POLICY_DECL(5, 13, 5)
BALLENV_LEAN_POLICY_DECL(10, 13, 5)
BALLENV_LEAN_POLICY_DECL(10, 8, 24)
BALLENV_LEAN_POLICY_DECL(5, 8, 24)
#undef BALLENV_LEAN_DECL
}  // namespace ballenv

// ballenv_state_written: is every coordinate of the (fp32) state integral and inside the ranges the lean kernels'
// exact shortcuts assume (ballenv_lean.cuh: small_integral for obstacles, small_int for agent and goal)?
__global__ void __launch_bounds__(128) validate_state_kernel(const __grid_constant__ ballenv::Params p, uint32_t* dirty) {
  using namespace ballenv;
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= p.n) return;
  const float* ax = reinterpret_cast<const float*>(p.agent_x);
  const float* ay = reinterpret_cast<const float*>(p.agent_y);
  const float* gx = reinterpret_cast<const float*>(p.goal_x);
  const float* gy = reinterpret_cast<const float*>(p.goal_y);
  bool ok = small_int(ax[e]) && small_int(ay[e]) && small_int(gx[e]) && small_int(gy[e]);
  const float* sx = reinterpret_cast<const float*>(p.stat_x) + e * p.stat_stride;
  const float* sy = reinterpret_cast<const float*>(p.stat_y) + e * p.stat_stride;
  for (int k = 0; k < p.cfg.ks; ++k) ok = ok && lean::small_integral(sx[k]) && lean::small_integral(sy[k]);
  const float* dx = reinterpret_cast<const float*>(p.dyn_x) + e * p.dyn_stride;
  const float* dy = reinterpret_cast<const float*>(p.dyn_y) + e * p.dyn_stride;
  for (int k = 0; k < p.cfg.kd; ++k) ok = ok && lean::small_integral(dx[k]) && lean::small_integral(dy[k]);
  if (!ok) atomicOr(dirty, 1u);
}

__global__ void selftest_sqrt_kernel(long long n, unsigned long long* bad) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float s = (float)i;
  if (__double_as_longlong(ballenv::sqrt_int22(s)) != __double_as_longlong(sqrt((double)s))) atomicAdd(bad, 1ull);
}

// the lean kernels' inline division (ballenv_lean.cuh: div64_fast_path) against the compiler's `/` on pairs shaped like
// the reward's: num = difference of two distances of integral points (|num| <= 1.5, often tiny), den = a distance in
// [471, 708] (total_distance of a gym-ruleset episode)
__global__ void selftest_div_kernel(long long n, unsigned long long* bad) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint4 w = ballenv::philox4x32_10((uint32_t)i, (uint32_t)(i >> 32), 0u, 77u, 0x1234u, 0x5678u);
  // two integral points at most one step apart, a goal and a start point: num = |p0 - g| - |p1 - g|, den = |s - g|
  const double gx = (double)(w.x % 500u), gy = 480.0 + (double)(w.y % 20u);
  const double px = (double)(w.z % 501u), py = (double)((w.z >> 16) % 501u);
  const double qx = px + (double)((int)(w.w % 3u) - 1), qy = py + (double)((int)((w.w >> 8) % 3u) - 1);
  const double sx = (double)((w.w >> 16) % 500u), sy = (double)((w.y >> 16) % 10u);
  const double num = ballenv::dist64(gx, gy, px, py) - ballenv::dist64(gx, gy, qx, qy);
  const double den = ballenv::dist64(gx, gy, sx, sy);
  if (num == 0.0) return;
  if (__double_as_longlong(ballenv::lean::div64_fast_path(num, den)) != __double_as_longlong(num / den)) atomicAdd(bad, 1ull);
}

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define CUDA_TRY(expr)                                                                     \
  do {                                                                                     \
    cudaError_t err__ = (expr);                                                            \
    if (err__ != cudaSuccess)                                                              \
      return fail(BALLENV_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__), __FILE__, __LINE__); \
  } while (0)

struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
};

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int obs_row_elems(const BallenvConfig& c) {
  const int nb = 4 + c.window * c.window;
  switch (c.obs_format) {
    case BALLENV_OBS_F32:
    case BALLENV_OBS_U8: return nb;
    case BALLENV_OBS_BITS: return (nb + 31) / 32;
  }
  return 0;
}

size_t obs_elem_bytes(const BallenvConfig& c) { return c.obs_format == BALLENV_OBS_U8 ? 1 : 4; }

int action_bytes(int kind) {
  switch (kind) {
    case BALLENV_ACT_INDEX_I64: return 8;
    case BALLENV_ACT_INDEX_I32: return 4;
    case BALLENV_ACT_INDEX_U8: return 1;
    case BALLENV_ACT_XY_F32: return 8;
    case BALLENV_ACT_XY_F64: return 16;
  }
  return 0;
}

int validate(const BallenvConfig* c) {
  if (c == nullptr) return fail(BALLENV_EINVAL, "config is NULL");
  if (c->abi_version != BALLENV_ABI_VERSION)
    return fail(BALLENV_EINVAL, "config.abi_version %d != %d", c->abi_version, BALLENV_ABI_VERSION);
  if (c->ruleset != BALLENV_RULESET_GYM && c->ruleset != BALLENV_RULESET_PYGAME)
    return fail(BALLENV_EINVAL, "unknown ruleset %d", c->ruleset);
  if (c->window < 1 || c->window > BALLENV_MAX_WINDOW)
    return fail(BALLENV_EINVAL, "window %d outside [1, %d]", c->window, BALLENV_MAX_WINDOW);
  if (c->static_obstacles < 0 || c->static_obstacles > BALLENV_MAX_STATIC)
    return fail(BALLENV_EINVAL, "static_obstacles %d outside [0, %d]", c->static_obstacles, BALLENV_MAX_STATIC);
  if (c->dynamic_obstacles < 0 || c->dynamic_obstacles > BALLENV_MAX_DYNAMIC)
    return fail(BALLENV_EINVAL, "dynamic_obstacles %d outside [0, %d]", c->dynamic_obstacles, BALLENV_MAX_DYNAMIC);
  if (c->precision != BALLENV_F32 && c->precision != BALLENV_F64)
    return fail(BALLENV_EINVAL, "unknown precision %d", c->precision);
  if (c->obs_format < BALLENV_OBS_F32 || c->obs_format > BALLENV_OBS_BITS)
    return fail(BALLENV_EINVAL, "unsupported obs_format %d", c->obs_format);
  if (c->max_episode_steps < 0) return fail(BALLENV_EINVAL, "max_episode_steps < 0");
  if (c->ruleset == BALLENV_RULESET_PYGAME) {
    // createBoard's dynamic-obstacle branch references undefined names (ballenv_pygame.py:502-506): unusable there too.
    if (c->dynamic_obstacles != 0)
      return fail(BALLENV_EINVAL, "the pygame ruleset has no working dynamic obstacles (ballenv_pygame.py:502-506)");
    if (!(c->agent_radius >= 0) || !(c->static_obstacle_radius >= 0))
      return fail(BALLENV_EINVAL, "negative radius");
  } else if (c->dynamic_obstacles > 0) {
    if (c->n_goals < c->dynamic_obstacles || c->n_goals > BALLENV_MAX_GOALS)
      return fail(BALLENV_EINVAL, "n_goals %d must be in [dynamic_obstacles = %d, %d] (ballenv_env.py:160)",
                  c->n_goals, c->dynamic_obstacles, BALLENV_MAX_GOALS);
    if (c->n_goals - 1 == 100)
      return fail(BALLENV_EINVAL, "101 obstacle goals are not supported");
    // every goal needs at least one *different* goal to change to, else np.random.randint(0) raises
    // at the first change step (ballenv_env.py:351-352)
    for (int i = 0; i < c->n_goals; ++i) {
      bool other = false;
      for (int k = 0; k < c->n_goals; ++k)
        other |= c->obs_goal_x[k] != c->obs_goal_x[i] || c->obs_goal_y[k] != c->obs_goal_y[i];
      if (!other) return fail(BALLENV_EINVAL, "obstacle goals must contain at least two distinct positions");
    }
    if (c->time_step_for_change < 0 || c->time_step_for_change >= (1 << 23))
      return fail(BALLENV_EINVAL, "time_step_for_change out of range");
  }
  return BALLENV_OK;
}

// The configuration a handle runs with.  The pygame ruleset always stores and computes in fp64: its coordinates are
// non-integral (uniform doubles, raw float actions) and its progress reward is the difference of two nearby distances,
// so fp32 positions cannot hold the 1e-5 relative reward tolerance; it is off the performance path anyway.
BallenvConfig effective_config(const BallenvConfig& c) {
  BallenvConfig e = c;
  if (e.ruleset == BALLENV_RULESET_PYGAME) e.precision = BALLENV_F64;
  return e;
}

struct Layout {
  size_t agent_x, agent_y, goal_x, goal_y, dist, total, acc, ep_len, episode, tick;
  size_t stat_x, stat_y, dyn_x, dyn_y, dyn_meta, flags, stats, errors, lean_tab, lean_tab_entries, bytes;
  long long stride, stat_stride, dyn_stride;
};

Layout make_layout(const BallenvConfig& c, long long n) {
  Layout L{};
  const size_t rb = c.precision == BALLENV_F64 ? 8 : 4;
  const long long S = (long long)align_up((size_t)n, 128);
  L.stride = S;
  L.stat_stride = (long long)align_up((size_t)c.static_obstacles, 4);   // environment-major rows of 128-bit quads
  L.dyn_stride = (long long)align_up((size_t)c.dynamic_obstacles, 4);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t at = off;
    off = align_up(off + bytes, 256);
    return at;
  };
  L.agent_x = take(rb * S);
  L.agent_y = take(rb * S);
  L.goal_x = take(rb * S);
  L.goal_y = take(rb * S);
  L.dist = take(8 * S);
  L.total = take(8 * S);
  L.acc = take(8 * S);
  L.ep_len = take(4 * S);
  L.episode = take(4 * S);
  L.tick = take(4 * S);
  L.stat_x = take(rb * S * L.stat_stride);
  L.stat_y = take(rb * S * L.stat_stride);
  L.dyn_x = take(rb * S * L.dyn_stride);
  L.dyn_y = take(rb * S * L.dyn_stride);
  L.dyn_meta = take(4 * S * L.dyn_stride);
  L.flags = take(S);
  L.stats = take(8 * BALLENV_NUM_STATS);
  L.errors = take(256);
  // column-mask table of the thread-per-environment kernels (ballenv_lean.cuh: LeanTab<W>), gym ruleset only
  L.lean_tab_entries = 0;
  if (c.ruleset == BALLENV_RULESET_GYM && c.window > 1 && c.window <= 16) {
    const int m = 25 + c.window / 2 + 2;
    L.lean_tab_entries = (size_t)(2 * m + 1) * (size_t)(2 * m + c.window - 1);
  }
  L.lean_tab = take(2 * L.lean_tab_entries + 16);
  L.bytes = off;
  return L;
}

}  // namespace

struct BallenvHandle {
  BallenvConfig cfg;
  long long n = 0, g0 = 0;
  int device = 0;
  uint64_t seed = 0;
  char* arena = nullptr;
  bool owns_arena = false;
  Layout L{};
  Params base{};
  // draw tapes (device copies)
  uint32_t* step_tape = nullptr;
  long long step_tape_steps = 0, step_tape_pos = 0;
  uint32_t* reset_tape = nullptr;
  // device staging for ballenv_step_host
  char* stage = nullptr;
  size_t stage_bytes = 0;
  size_t stage_act = 0, stage_obs = 0, stage_rew = 0, stage_done = 0;
  // ballenv_step_many_host: copy streams either way and the events that order them against the compute stream
  char* many = nullptr;         // chunk staging of ballenv_step_many_host (two sets)
  size_t many_bytes = 0;
  cudaStream_t s_in = nullptr, s_out = nullptr;
  cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_k[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr}, ev_start = nullptr;
  // ballenv_observe_patches: device tables of the last (width, out_size, interp) asked for
  char* patch_tab = nullptr;
  int patch_key[3] = {0, 0, -1};
  int patch_ksize = 0;
  long long launches = 0;
  bool no_rollout = false;      // BALLENV_NO_ROLLOUT=1: ballenv_step_many launches one kernel per step (tests, profiling)
  int lean_g = 0;               // BALLENV_LEAN_G=1|2: lanes per environment of the lean kernels (0: the measured best)
  bool no_lean = false;         // BALLENV_NO_LEAN=1: never pick the thread-per-environment kernels (tests, A/B runs)
  bool lean_rt_only = false;    // BALLENV_LEAN_RT=1: the run-time-count lean kernels even where a fixed instance exists (tests, A/B runs)
  bool force_generic = false;   // BALLENV_FORCE_GENERIC=1 in the environment: never pick the fast specialisation (tests)
};

namespace {

// production specialisation (see ballenv_kernel<.., kFast>): everything the generic kernel tests per launch
// any_rows: the observation rows may have any format (the lean kernels); otherwise float32 rows (block of roles)
bool fast_eligible(const BallenvHandle* h, const Params& p, bool any_rows = false) {
  return h->cfg.precision == BALLENV_F32 && p.mode == kModeStep && p.cfg.ruleset == BALLENV_RULESET_GYM &&
         p.step_tape == nullptr && p.reset_tape == nullptr && (p.cfg.goals_distinct || p.cfg.kd == 0) &&
         p.obs != nullptr && (any_rows || p.cfg.obs_format == BALLENV_OBS_F32) &&
         (p.action_kind == BALLENV_ACT_INDEX_I64 || p.action_kind == BALLENV_ACT_INDEX_I32 ||
          p.action_kind == BALLENV_ACT_INDEX_U8) &&
         !h->force_generic;
}

// thread-per-environment kernels (ballenv_lean.cuh): the production configuration with one of the instantiated
// obstacle counts, integral geometry and a change step that fits a byte
typedef void (*LeanLauncher)(const Params&, unsigned, cudaStream_t);
LeanLauncher lean_launcher(const BallenvHandle* h, const Params& p, int* lanes = nullptr) {
  if (!fast_eligible(h, p, true) || h->no_lean || p.lean_tab == nullptr) return nullptr;
  const DevConfig& c = p.cfg;
  if (c.change_step > 254 || c.n_goals < 2 || c.step_x != 1.0 || c.step_y != 1.0 || c.radius_sum != 25.0) return nullptr;
  if (c.margin != (double)(25 + c.window / 2 + 2)) return nullptr;
  // g = lanes per environment the configuration runs best with (measured, tools/lean_ab.sh): rollouts - a pair for the
  // 32-obstacle configurations (C3: 6.5 us per step against 8.2), one lane for the reference's default 13 + 5 (3.3
  // against 3.7); single-step launches - always a pair (13 + 5: 12.2 us per launch against 15.2: the launch is a
  // chain of latencies, and a pair halves every lane's share of it)
  struct Inst { int w, ks, kd, g; LeanLauncher g1, g2; };
  static const Inst kInst[] = {{5, 13, 5, 1, launch_lean_w5_s13_d5_g1, launch_lean_w5_s13_d5_g2},
                               {10, 13, 5, 1, launch_lean_w10_s13_d5_g1, launch_lean_w10_s13_d5_g2},
                               {10, 8, 24, 2, launch_lean_w10_s8_d24_g1, launch_lean_w10_s8_d24_g2},
                               {5, 8, 24, 2, launch_lean_w5_s8_d24_g1, launch_lean_w5_s8_d24_g2}};
  for (const Inst& i : kInst)
    if (i.w == c.window && i.ks == c.ks && i.kd == c.kd && !h->lean_rt_only) {
      const int g = h->lean_g ? h->lean_g : (p.n_steps > 1 ? i.g : 2);
      if (lanes != nullptr) *lanes = g;
      return g == 2 ? i.g2 : i.g1;
    }
  // any other obstacle counts: the same kernels reading the counts at run time (ballenv_lean_rt.cuh), as long as the
  // lists fit (the reset's near mask holds 64 obstacles) and at least two blocks' regions fit an SM's shared memory.
  // Measured against the block-of-roles kernel at 65 536 environments (tools/rt_rate.py): single-step launches - a pair
  // of lanes, always (13 + 5 counts 10.1 us against 15.0, 8 + 24 counts 18.1 against 22.1); rollouts - one lane per
  // environment while the moving obstacles are few (13 + 5: 4.0 us per step against 5.9), the block of roles from four
  // moving quads up (8 + 24: 11.1 against 9.1: the run-time form's larger shared-memory region leaves five blocks per
  // SM where the launch needs seven to run in one wave - tools/rt_ncu.sh).
  if ((c.window == 5 || c.window == 10) && c.kd >= 1 && c.ks + c.kd <= kLeanRtMaxObstacles) {
    const int dyn_quads = (c.kd + 3) / 4;
    if (p.n_steps > 1 && dyn_quads >= 4 && !h->lean_g && !h->lean_rt_only) return nullptr;
    const int g = h->lean_g ? h->lean_g : (p.n_steps > 1 ? 1 : 2);
    const size_t need = 4096 + (c.window == 5 ? (g == 2 ? lean_rt_smem_w5_g2(c.ks, c.kd) : lean_rt_smem_w5_g1(c.ks, c.kd))
                                              : (g == 2 ? lean_rt_smem_w10_g2(c.ks, c.kd) : lean_rt_smem_w10_g1(c.ks, c.kd)));
    if (2 * need <= 220 * 1024) {
      if (lanes != nullptr) *lanes = g;
      if (c.window == 5) return g == 2 ? launch_lean_w5_rt_g2 : launch_lean_w5_rt_g1;
      return g == 2 ? launch_lean_w10_rt_g2 : launch_lean_w10_rt_g1;
    }
  }
  return nullptr;
}

// the lean kernel with the policy in its loop for this configuration, or nullptr: a tuned instance, else the
// run-time-count form (any counts up to 64 obstacles) when its regions and the policy fit shared memory together
LeanLauncher policy_launcher(const BallenvHandle* h, const Params& p, int* lanes = nullptr) {
  if (!fast_eligible(h, p, true) || h->no_lean || p.lean_tab == nullptr || p.cfg.obs_format != BALLENV_OBS_F32) return nullptr;
  const DevConfig& c = p.cfg;
  if (c.change_step > 254 || c.n_goals < 2 || c.step_x != 1.0 || c.step_y != 1.0 || c.radius_sum != 25.0) return nullptr;
  if (c.margin != (double)(25 + c.window / 2 + 2)) return nullptr;
  struct Inst { int w, ks, kd; LeanLauncher g1, g2; };
  static const Inst kInst[] = {{5, 13, 5, launch_lean_policy_w5_s13_d5_g1, launch_lean_policy_w5_s13_d5_g2},
                               {10, 13, 5, launch_lean_policy_w10_s13_d5_g1, launch_lean_policy_w10_s13_d5_g2},
                               {10, 8, 24, launch_lean_policy_w10_s8_d24_g1, launch_lean_policy_w10_s8_d24_g2},
                               {5, 8, 24, launch_lean_policy_w5_s8_d24_g1, launch_lean_policy_w5_s8_d24_g2}};
  // a pair of lanes per environment unless told otherwise: the pair also splits the hidden units of the policy, the
  // longest dependent chain of a step
  const int g = h->lean_g ? h->lean_g : 2;
  if (lanes != nullptr) *lanes = g;
  if (!h->lean_rt_only)
    for (const Inst& i : kInst)
      if (i.w == c.window && i.ks == c.ks && i.kd == c.kd) return g == 2 ? i.g2 : i.g1;
  if ((c.window == 5 || c.window == 10) && c.kd >= 1 && c.ks + c.kd <= kLeanRtMaxObstacles) {
    const size_t regions = c.window == 5 ? (g == 2 ? lean_rt_smem_w5_g2(c.ks, c.kd) : lean_rt_smem_w5_g1(c.ks, c.kd))
                                         : (g == 2 ? lean_rt_smem_w10_g2(c.ks, c.kd) : lean_rt_smem_w10_g1(c.ks, c.kd));
    if (regions + 4 * lean::policy_smem_floats(4 + c.window * c.window, p.pol_hidden) + 4096 <= 220 * 1024) {
      if (c.window == 5) return g == 2 ? launch_lean_policy_w5_rt_g2 : launch_lean_policy_w5_rt_g1;
      return g == 2 ? launch_lean_policy_w10_rt_g2 : launch_lean_policy_w10_rt_g1;
    }
  }
  return nullptr;
}

int launch(BallenvHandle* h, const Params& p_in, cudaStream_t s) {
  Params p = p_in;   // + the constants the kernels would otherwise derive per step
  p.obs_row_bytes = (long long)p.cfg.obs_row_elems * (p.cfg.obs_format == BALLENV_OBS_U8 ? 1 : 4);
  p.obs_step_bytes = (p.n_steps > 1 && p.obs_all_steps) ? p.n * p.obs_row_bytes : 0;
  {
    // obstacle threads of a block: one per quad.  In the fast single-step kernel a static-quad thread takes two
    // quads when that makes the block fit 4 obstacle warps (kernel parameter kSQ = 2: seven resident blocks; measured
    // 16.4 -> 15.0 us per launch for the reference's default 13 + 5 obstacles).  Not for rollouts, and not for C3
    // (one static warp instead of two becomes the long pole of the step: 9.1 -> 10.2 us, measured).
    const int qs = (p.cfg.ks + 3) / 4, qd = (p.cfg.kd + 3) / 4;
    p.sq = (fast_eligible(h, p) && p.n_steps == 1 && (qs == 2 || qs == 4) && qs / 2 + qd <= 4) ? 2 : 1;
    p.n_stat = kEnvsPerBlock * (qs / p.sq);
    p.n_slot = p.n_stat + kEnvsPerBlock * qd;
  }
  if (LeanLauncher lean = lean_launcher(h, p)) {
    lean(p, (unsigned)((p.n + kLeanEnvsPerBlock - 1) / kLeanEnvsPerBlock), s);
    h->launches += 1;
    CUDA_TRY(cudaGetLastError());
    return BALLENV_OK;
  }
  const unsigned grid = (unsigned)((p.n + kEnvsPerBlock - 1) / kEnvsPerBlock);
  const bool f64 = h->cfg.precision == BALLENV_F64;
  const bool fast = fast_eligible(h, p);
  if (!fast && p.n_steps != 1) return fail(BALLENV_EINVAL, "internal: multi-step launch needs a rollout kernel");
  if (fast) {
    switch (h->cfg.window) {
      case 5: launch_f32_w5_fast(p, grid, s); break;
      case 10: launch_f32_w10_fast(p, grid, s); break;
      default: launch_f32_wany_fast(p, grid, s); break;
    }
    h->launches += 1;
    CUDA_TRY(cudaGetLastError());
    return BALLENV_OK;
  }
  switch (h->cfg.window) {
    case 5: f64 ? launch_f64_w5(p, grid, s) : launch_f32_w5(p, grid, s); break;
    case 10: f64 ? launch_f64_w10(p, grid, s) : launch_f32_w10(p, grid, s); break;
    default: f64 ? launch_f64_wany(p, grid, s) : launch_f32_wany(p, grid, s); break;
  }
  h->launches += 1;
  CUDA_TRY(cudaGetLastError());
  return BALLENV_OK;
}

void fill_dev_config(const BallenvConfig& c, DevConfig* d) {
  memset(d, 0, sizeof(*d));
  d->ruleset = c.ruleset;
  d->window = c.window;
  d->ks = c.static_obstacles;
  d->kd = c.dynamic_obstacles;
  d->n_goals = c.dynamic_obstacles > 0 ? c.n_goals : 0;
  d->change_step = c.time_step_for_change;
  d->rd_th = c.rd_th_obs;
  d->max_steps = c.max_episode_steps;
  d->auto_reset = c.auto_reset;
  d->obs_format = c.obs_format;
  d->obs_row_elems = obs_row_elems(c);
  d->static_penalty = c.static_penalty;
  d->dynamic_penalty = c.dynamic_penalty;
  if (c.ruleset == BALLENV_RULESET_GYM) {
    d->world_w = 500.0;       // _screen_width / _screen_height, ballenv_env.py:11-12
    d->world_h = 500.0;
    d->radius_sum = 20.0 + 5.0;   // radius_rand_person + radius_ctrl_person, :49-50,188
    d->goal_threshold = 10.0;     // :64
    d->step_x = 1.0;              // speedx_ctrl_person / speedy_ctrl_person, :53-54
    d->step_y = 1.0;
  } else {
    d->world_w = 100.0;       // ballenv_pygame.py:8-9
    d->world_h = 100.0;
    d->radius_sum = c.static_obstacle_radius + c.agent_radius;   // :384
    d->goal_threshold = 15.0;     // :345
    d->step_x = 1.0;
    d->step_y = 1.0;
    d->reset_agent_thresh = 15.0; // :494
    d->reset_goal_thresh = 5.0;
  }
  d->margin = d->radius_sum + (d->step_x > d->step_y ? d->step_x : d->step_y) * (double)(c.window / 2) + 2.0;
  d->f_world_w = (float)d->world_w;
  d->f_world_h = (float)d->world_h;
  d->f_step_x = (float)d->step_x;
  d->f_step_y = (float)d->step_y;
  d->f_margin = (float)d->margin;
  const uint32_t qs = (uint32_t)(c.static_obstacles + 3) / 4, qd = (uint32_t)(c.dynamic_obstacles + 3) / 4;
  d->rcp_qs = qs > 1 ? (uint32_t)(((1ull << 32) + qs - 1) / qs) : 0u;
  d->rcp_qd = qd > 1 ? (uint32_t)(((1ull << 32) + qd - 1) / qd) : 0u;
  bool distinct = true;
  for (int i = 0; i < d->n_goals; ++i) {
    d->goal_x[i] = c.obs_goal_x[i];
    d->goal_y[i] = c.obs_goal_y[i];
    for (int k = 0; k < i; ++k)
      distinct &= c.obs_goal_x[k] != c.obs_goal_x[i] || c.obs_goal_y[k] != c.obs_goal_y[i];
  }
  d->goals_distinct = distinct ? 1 : 0;
  for (int i = 0; i < d->n_goals; ++i) d->f_goal[i] = make_float2((float)c.obs_goal_x[i], (float)c.obs_goal_y[i]);
  d->lean_integral_speeds = 1;
  d->lean_cs4 = (uint32_t)(c.time_step_for_change & 0xff) * 0x01010101u;
  for (int j = 0; j < c.dynamic_obstacles; ++j) {
    d->speed[j] = c.obstacle_speed[j];
    d->f_speed[j] = (float)c.obstacle_speed[j];
    if (c.obstacle_speed[j] != (double)(long long)c.obstacle_speed[j] || c.obstacle_speed[j] > 1024.0 ||
        c.obstacle_speed[j] < -1024.0)
      d->lean_integral_speeds = 0;
  }
}

int ensure_stage(BallenvHandle* h, int action_kind) {
  const size_t n = (size_t)h->n;
  const size_t a = align_up(n * 16, 256);   // largest action layout
  const size_t o = align_up(n * (size_t)obs_row_elems(h->cfg) * obs_elem_bytes(h->cfg), 256);
  const size_t r = align_up(n * 8, 256);
  const size_t d = align_up(n, 256);
  const size_t need = a + o + r + d;
  (void)action_kind;
  if (h->stage_bytes < need) {
    if (h->stage) cudaFree(h->stage);
    h->stage = nullptr;
    h->stage_bytes = 0;
    CUDA_TRY(cudaMalloc(&h->stage, need));
    h->stage_bytes = need;
    h->stage_act = 0;
    h->stage_obs = a;
    h->stage_rew = a + o;
    h->stage_done = a + o + r;
  }
  return BALLENV_OK;
}

}  // namespace

extern "C" {

int ballenv_abi_version(void) { return BALLENV_ABI_VERSION; }

const char* ballenv_last_error(void) { return g_err; }

int ballenv_config_default(BallenvConfig* cfg, int ruleset) {
  if (cfg == nullptr) return fail(BALLENV_EINVAL, "cfg is NULL");
  memset(cfg, 0, sizeof(*cfg));
  cfg->abi_version = BALLENV_ABI_VERSION;
  cfg->ruleset = ruleset;
  cfg->window = 5;
  cfg->max_episode_steps = 1000;
  cfg->auto_reset = 1;
  cfg->precision = BALLENV_F32;
  cfg->obs_format = BALLENV_OBS_F32;
  cfg->agent_radius = 10.0;
  cfg->static_obstacle_radius = 10.0;
  if (ruleset == BALLENV_RULESET_GYM) {  // examples/ball_cnn_ac3.py:40-51
    static const double gx[5] = {12, 123, 87, 430, 230}, gy[5] = {122, 93, 150, 440, 11};
    cfg->static_obstacles = 13;
    cfg->dynamic_obstacles = 5;
    cfg->n_goals = 5;
    for (int i = 0; i < 5; ++i) {
      cfg->obstacle_speed[i] = 1.0;
      cfg->obs_goal_x[i] = gx[i];
      cfg->obs_goal_y[i] = gy[i];
    }
    cfg->time_step_for_change = 50;
    cfg->rd_th_obs = 60;
    cfg->static_penalty = 1.0;
    cfg->dynamic_penalty = 8000.0;
  } else if (ruleset == BALLENV_RULESET_PYGAME) {  // createBoard() defaults, ballenv_pygame.py:316
    cfg->static_obstacles = 0;
    cfg->max_episode_steps = 0;
  } else {
    return fail(BALLENV_EINVAL, "unknown ruleset %d", ruleset);
  }
  return BALLENV_OK;
}

int64_t ballenv_state_bytes(const BallenvConfig* cfg, int64_t n_envs) {
  int rc = validate(cfg);
  if (rc != BALLENV_OK) return rc;
  if (n_envs <= 0) return fail(BALLENV_EINVAL, "n_envs must be positive");
  return (int64_t)make_layout(effective_config(*cfg), n_envs).bytes;
}

int ballenv_create(const BallenvConfig* cfg, int64_t n_envs, int64_t global_env_offset, int device, uint64_t seed,
                   void* arena, BallenvHandle** out) {
  if (out == nullptr) return fail(BALLENV_EINVAL, "out is NULL");
  *out = nullptr;
  int rc = validate(cfg);
  if (rc != BALLENV_OK) return rc;
  if (n_envs <= 0 || n_envs > (1ll << 31) - 256) return fail(BALLENV_EINVAL, "n_envs %lld out of range", (long long)n_envs);
  if (global_env_offset < 0 || global_env_offset + n_envs > (1ll << 32))
    return fail(BALLENV_EINVAL, "global env ids must fit 32 bits");
  {
    // the kernels address obstacle quads with 32-bit element offsets: n x (obstacles rounded up to 4) must fit
    const long long kmax = (long long)align_up((size_t)(cfg->static_obstacles > cfg->dynamic_obstacles
                                                             ? cfg->static_obstacles : cfg->dynamic_obstacles), 4);
    if ((long long)align_up((size_t)n_envs, 128) * (kmax > 0 ? kmax : 1) >= (1ll << 31))
      return fail(BALLENV_EINVAL, "n_envs x obstacles per kind (%lld x %lld) must stay below 2^31 elements",
                  (long long)n_envs, kmax);
  }
  int ndev = 0;
  CUDA_TRY(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(BALLENV_EINVAL, "device %d not in [0, %d)", device, ndev);
  DeviceGuard guard(device);
  BallenvHandle* h = new (std::nothrow) BallenvHandle();
  if (h == nullptr) return fail(BALLENV_ENOMEM, "out of host memory");
  h->cfg = effective_config(*cfg);
  h->n = n_envs;
  h->g0 = global_env_offset;
  h->device = device;
  h->seed = seed;
  h->L = make_layout(h->cfg, n_envs);
  const char* fg = getenv("BALLENV_FORCE_GENERIC");
  h->force_generic = fg != nullptr && fg[0] == '1';
  const char* nl = getenv("BALLENV_NO_LEAN");
  h->no_lean = nl != nullptr && nl[0] == '1';
  const char* lr = getenv("BALLENV_LEAN_RT");
  h->lean_rt_only = lr != nullptr && lr[0] == '1';
  const char* lg = getenv("BALLENV_LEAN_G");
  h->lean_g = lg != nullptr && (lg[0] == '1' || lg[0] == '2') ? lg[0] - '0' : 0;
  const char* nr = getenv("BALLENV_NO_ROLLOUT");
  h->no_rollout = nr != nullptr && nr[0] == '1';
  if (arena != nullptr) {
    if (((uintptr_t)arena & 255) != 0) {
      delete h;
      return fail(BALLENV_EINVAL, "arena must be 256-byte aligned");
    }
    h->arena = (char*)arena;
  } else {
    cudaError_t e = cudaMalloc(&h->arena, h->L.bytes);
    if (e != cudaSuccess) {
      delete h;
      return fail(BALLENV_ENOMEM, "cudaMalloc(%zu) failed: %s", h->L.bytes, cudaGetErrorString(e));
    }
    h->owns_arena = true;
  }
  cudaError_t e = cudaMemset(h->arena, 0, h->L.bytes);
  if (e == cudaSuccess) e = cudaMemset(h->arena + h->L.episode, 0xff, 4 * (size_t)h->L.stride);  // episode = -1: none yet
  if (e == cudaSuccess && h->L.lean_tab_entries > 0) {
    // column masks of the exact raster (LeanTab<W> in ballenv_lean.cuh): entry [ui][s] = columns c of the window with
    // (c - u)^2 + dv^2 <= 25^2, u = ui + h - M, dv = s - h - M (examples/ball_cnn_ac3.py:396-409 for integral coordinates)
    const int w = cfg->window, hh = w / 2, m = 25 + hh + 2, U = 2 * m + 1, S = 2 * m + w - 1;
    uint16_t* tab = (uint16_t*)malloc(sizeof(uint16_t) * (size_t)U * S);
    if (tab == nullptr) e = cudaErrorMemoryAllocation;
    else {
      for (int ui = 0; ui < U; ++ui)
        for (int si = 0; si < S; ++si) {
          const int u = ui + hh - m, dv = si - hh - m;
          uint16_t mask = 0;
          for (int cc = 0; cc < w; ++cc)
            if ((cc - u) * (cc - u) + dv * dv <= 625) mask |= (uint16_t)(1u << cc);
          tab[(size_t)ui * S + si] = mask;
        }
      e = cudaMemcpy(h->arena + h->L.lean_tab, tab, sizeof(uint16_t) * (size_t)U * S, cudaMemcpyHostToDevice);
      free(tab);
    }
  }
  if (e == cudaSuccess) e = cudaMemset(h->arena + h->L.errors + 4, 1, 4);   // state_dirty: nothing is known yet
  // everything above ran on the legacy stream: a caller that launches on a non-blocking stream next must see it done
  if (e == cudaSuccess) e = cudaStreamSynchronize(0);
  if (e != cudaSuccess) {
    if (h->owns_arena) cudaFree(h->arena);
    delete h;
    return fail(BALLENV_ECUDA, "arena initialisation failed: %s", cudaGetErrorString(e));
  }
  Params& p = h->base;
  memset(&p, 0, sizeof(p));
  fill_dev_config(h->cfg, &p.cfg);
  const Layout& L = h->L;
  p.n = n_envs;
  p.stride = L.stride;
  p.stat_stride = L.stat_stride;
  p.dyn_stride = L.dyn_stride;
  p.g0 = (uint32_t)global_env_offset;
  p.k0 = (uint32_t)(seed & 0xffffffffu);
  p.k1 = (uint32_t)(seed >> 32);
  const char* dbg = getenv("BALLENV_DEBUG_SKIP");
  p.debug = dbg != nullptr ? atoi(dbg) : 0;
  char* a = h->arena;
  p.agent_x = a + L.agent_x;
  p.agent_y = a + L.agent_y;
  p.goal_x = a + L.goal_x;
  p.goal_y = a + L.goal_y;
  p.dist = (double*)(a + L.dist);
  p.total = (double*)(a + L.total);
  p.acc = (double*)(a + L.acc);
  p.ep_len = (int*)(a + L.ep_len);
  p.episode = (uint32_t*)(a + L.episode);
  p.tick = (uint32_t*)(a + L.tick);
  p.stat_x = a + L.stat_x;
  p.stat_y = a + L.stat_y;
  p.dyn_x = a + L.dyn_x;
  p.dyn_y = a + L.dyn_y;
  p.dyn_meta = (uint32_t*)(a + L.dyn_meta);
  p.flags = (uint8_t*)(a + L.flags);
  p.stats = (double*)(a + L.stats);
  p.errors = (uint32_t*)(a + L.errors);
  p.lean_tab = L.lean_tab_entries > 0 ? (const uint16_t*)(a + L.lean_tab) : nullptr;
  p.state_dirty = p.errors + 1;   // second word of the error block: "coordinates not known to be integral"
  for (int r = 0; r < 10; ++r) {   // Philox4x32 key schedule (ballenv_rng.cuh)
    p.rk[2 * r] = p.k0 + (uint32_t)r * kPhiloxW0;
    p.rk[2 * r + 1] = p.k1 + (uint32_t)r * kPhiloxW1;
  }
  *out = h;
  return BALLENV_OK;
}

int ballenv_destroy(BallenvHandle* h) {
  if (h == nullptr) return BALLENV_OK;
  DeviceGuard guard(h->device);
  cudaDeviceSynchronize();
  if (h->owns_arena && h->arena) cudaFree(h->arena);
  if (h->step_tape) cudaFree(h->step_tape);
  if (h->reset_tape) cudaFree(h->reset_tape);
  if (h->stage) cudaFree(h->stage);
  if (h->many) cudaFree(h->many);
  if (h->patch_tab) cudaFree(h->patch_tab);
  if (h->s_in) cudaStreamDestroy(h->s_in);
  if (h->s_out) cudaStreamDestroy(h->s_out);
  for (int i = 0; i < 2; ++i) {
    if (h->ev_in[i]) cudaEventDestroy(h->ev_in[i]);
    if (h->ev_k[i]) cudaEventDestroy(h->ev_k[i]);
    if (h->ev_out[i]) cudaEventDestroy(h->ev_out[i]);
  }
  if (h->ev_start) cudaEventDestroy(h->ev_start);
  delete h;
  return BALLENV_OK;
}

int ballenv_state_ptrs(BallenvHandle* h, BallenvStatePtrs* out) {
  if (h == nullptr || out == nullptr) return fail(BALLENV_EINVAL, "NULL argument");
  const Params& p = h->base;
  memset(out, 0, sizeof(*out));
  out->n_envs = h->n;
  out->n_stride = h->L.stride;
  out->static_stride = h->L.stat_stride;
  out->dynamic_stride = h->L.dyn_stride;
  out->real_bytes = h->cfg.precision == BALLENV_F64 ? 8 : 4;
  out->obs_row_elems = obs_row_elems(h->cfg);
  out->agent_x = p.agent_x;
  out->agent_y = p.agent_y;
  out->goal_x = p.goal_x;
  out->goal_y = p.goal_y;
  out->dist = p.dist;
  out->total_distance = p.total;
  out->acc_reward = p.acc;
  out->ep_len = p.ep_len;
  out->episode = p.episode;
  out->tick = p.tick;
  out->static_x = p.stat_x;
  out->static_y = p.stat_y;
  out->dynamic_x = p.dyn_x;
  out->dynamic_y = p.dyn_y;
  out->dynamic_meta = p.dyn_meta;
  out->flags = p.flags;
  out->stats = p.stats;
  out->error_flags = p.errors;
  return BALLENV_OK;
}

int ballenv_reset(BallenvHandle* h, const uint8_t* mask, void* obs_out, ballenv_stream_t stream) {
  if (h == nullptr) return fail(BALLENV_EINVAL, "handle is NULL");
  DeviceGuard guard(h->device);
  Params p = h->base;
  p.mode = kModeReset;
  p.n_steps = 1;
  p.reset_mask = mask;
  p.obs = obs_out;
  p.reset_tape = h->reset_tape;
  int rc = launch(h, p, (cudaStream_t)stream);
  // a full reset of the gym ruleset leaves nothing but integer draws in the state: the lean kernels may skip their test
  if (rc == BALLENV_OK && mask == nullptr && h->cfg.ruleset == BALLENV_RULESET_GYM && h->cfg.precision == BALLENV_F32)
    CUDA_TRY(cudaMemsetAsync(const_cast<uint32_t*>(h->base.state_dirty), 0, 4, (cudaStream_t)stream));
  return rc;
}

int ballenv_state_written(BallenvHandle* h, ballenv_stream_t stream) {
  if (h == nullptr) return fail(BALLENV_EINVAL, "handle is NULL");
  DeviceGuard guard(h->device);
  uint32_t* dirty = const_cast<uint32_t*>(h->base.state_dirty);
  if (h->cfg.precision != BALLENV_F32 || h->cfg.ruleset != BALLENV_RULESET_GYM) {   // no lean kernels there: stays unknown
    CUDA_TRY(cudaMemsetAsync(dirty, 1, 4, (cudaStream_t)stream));
    return BALLENV_OK;
  }
  CUDA_TRY(cudaMemsetAsync(dirty, 0, 4, (cudaStream_t)stream));
  validate_state_kernel<<<(unsigned)((h->n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(h->base, dirty);
  h->launches += 1;
  CUDA_TRY(cudaGetLastError());
  return BALLENV_OK;
}

int ballenv_reset_fixed(BallenvHandle* h, const uint8_t* mask, double goal_x, double goal_y, void* obs_out,
                        ballenv_stream_t stream) {
  if (h == nullptr) return fail(BALLENV_EINVAL, "handle is NULL");
  if (h->cfg.ruleset != BALLENV_RULESET_PYGAME)
    return fail(BALLENV_EINVAL, "resetFixedstate belongs to the pygame ruleset (ballenv_pygame.py:589-624)");
  DeviceGuard guard(h->device);
  const unsigned grid = (unsigned)((h->n + 127) / 128);
  ballenv_reset_fixed_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(h->base, mask, goal_x, goal_y);
  h->launches += 1;
  CUDA_TRY(cudaGetLastError());
  return obs_out != nullptr ? ballenv_observe(h, obs_out, stream) : BALLENV_OK;
}

int ballenv_observe(BallenvHandle* h, void* obs_out, ballenv_stream_t stream) {
  if (h == nullptr || obs_out == nullptr) return fail(BALLENV_EINVAL, "NULL argument");
  DeviceGuard guard(h->device);
  Params p = h->base;
  p.mode = kModeObserve;
  p.n_steps = 1;
  p.obs = obs_out;
  return launch(h, p, (cudaStream_t)stream);
}

int ballenv_observe_features(BallenvHandle* h, float* out, ballenv_stream_t stream) {
  if (h == nullptr || out == nullptr) return fail(BALLENV_EINVAL, "NULL argument");
  if (((uintptr_t)out & 15) != 0) return fail(BALLENV_EINVAL, "out must be 16-byte aligned");
  DeviceGuard guard(h->device);
  const unsigned grid = (unsigned)((h->n + 127) / 128);
  if (h->cfg.precision == BALLENV_F64)
    ballenv_features_kernel<double><<<grid, 128, 0, (cudaStream_t)stream>>>(h->base, out, h->cfg.agent_radius, 0.0, 0.0, 0.0, 0.0);
  else
    ballenv_features_kernel<float><<<grid, 128, 0, (cudaStream_t)stream>>>(h->base, out, h->cfg.agent_radius, 0.0, 0.0, 0.0, 0.0);
  h->launches += 1;
  CUDA_TRY(cudaGetLastError());
  return BALLENV_OK;
}

int ballenv_observe_blocks(BallenvHandle* h, float* out, ballenv_stream_t stream) {
  if (h == nullptr || out == nullptr) return fail(BALLENV_EINVAL, "NULL argument");
  DeviceGuard guard(h->device);
  const unsigned grid = (unsigned)((h->n + 127) / 128);
  if (h->cfg.precision == BALLENV_F64) ballenv_blocks_kernel<double><<<grid, 128, 0, (cudaStream_t)stream>>>(h->base, out);
  else ballenv_blocks_kernel<float><<<grid, 128, 0, (cudaStream_t)stream>>>(h->base, out);
  h->launches += 1;
  CUDA_TRY(cudaGetLastError());
  return BALLENV_OK;
}

int ballenv_observe_patches(BallenvHandle* h, void* out, int32_t width, int32_t out_size, int32_t interp,
                            int32_t out_format, ballenv_stream_t stream) {
  if (h == nullptr || out == nullptr) return fail(BALLENV_EINVAL, "NULL argument");
  if (h->cfg.ruleset != BALLENV_RULESET_GYM) return fail(BALLENV_EINVAL, "rgb patches render the gym ruleset's viewer");
  if (width < 2 || width > kPatchMaxWidth || (width & 1) != 0) return fail(BALLENV_EINVAL, "width must be even, 2..%d", kPatchMaxWidth);
  if (out_size < 1 || out_size > kPatchMaxOut) return fail(BALLENV_EINVAL, "out_size must be 1..%d", kPatchMaxOut);
  if (interp != BALLENV_INTERP_BILINEAR && interp != BALLENV_INTERP_BICUBIC) return fail(BALLENV_EINVAL, "unknown interp %d", interp);
  if (out_format != BALLENV_OBS_F32 && out_format != BALLENV_OBS_U8) return fail(BALLENV_EINVAL, "patches are float32 or uint8");
  DeviceGuard guard(h->device);
  cudaStream_t s = (cudaStream_t)stream;
  // device tables: [bounds 2 * 64][coef 64 * kmax][agent 10][goal 10][obstacle 40] - rebuilt when the geometry changes
  constexpr int kMaxK = 64;   // taps of the resampling kernel: ceil(support * width / out_size) * 2 + 1 must fit
  const size_t off_coef = 4 * 2 * kPatchMaxOut, off_masks = (off_coef + 4 * (size_t)kPatchMaxOut * kMaxK + 7) / 8 * 8;
  const size_t tab_bytes = off_masks + 8 * (2 * kPatchAgentR + 10 + 2 * kPatchObstacleR);
  if (h->patch_tab == nullptr) CUDA_TRY(cudaMalloc(&h->patch_tab, tab_bytes));
  if (h->patch_key[0] != width || h->patch_key[1] != out_size || h->patch_key[2] != interp) {
    std::vector<int> bounds, kk;
    const int ksize = patch_host::coefficients(width, out_size, interp == BALLENV_INTERP_BICUBIC, &bounds, &kk);
    if (ksize > kMaxK) return fail(BALLENV_EINVAL, "width / out_size too large: the resampling kernel has %d taps (max %d)", ksize, kMaxK);
    unsigned long long masks[2 * kPatchAgentR + 10 + 2 * kPatchObstacleR];
    patch_host::sprite_rows(patch_host::circle_polygon(kPatchAgentR), kPatchAgentR, masks);                       // ballenv_env.py:367
    patch_host::sprite_rows(std::vector<double>{5, 5, 5, -5, -5, 5, -5, -5}, 5, masks + 2 * kPatchAgentR);           // :371
    patch_host::sprite_rows(patch_host::circle_polygon(kPatchObstacleR), kPatchObstacleR, masks + 2 * kPatchAgentR + 10);   // :299, :305
    // (pageable sources: the copies are staged before these calls return)
    CUDA_TRY(cudaMemcpyAsync(h->patch_tab, bounds.data(), 4 * bounds.size(), cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(h->patch_tab + off_coef, kk.data(), 4 * kk.size(), cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(h->patch_tab + off_masks, masks, sizeof(masks), cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaStreamSynchronize(s));   // the host vectors go away
    h->patch_key[0] = width;
    h->patch_key[1] = out_size;
    h->patch_key[2] = interp;
    h->patch_ksize = ksize;
  }
  PatchTables tb;
  tb.width = width;
  tb.out = out_size;
  tb.ksize = h->patch_ksize;
  tb.bounds = reinterpret_cast<const int*>(h->patch_tab);
  tb.coef = reinterpret_cast<const int*>(h->patch_tab + off_coef);
  tb.agent = reinterpret_cast<const unsigned long long*>(h->patch_tab + off_masks);
  tb.goal = tb.agent + 2 * kPatchAgentR;
  tb.obstacle = tb.goal + 10;
  const int n_obj = 2 + h->cfg.static_obstacles + h->cfg.dynamic_obstacles;
  const size_t smem = patch_smem_bytes(width, out_size, tb.ksize, n_obj);
  if (smem > 200 * 1024) return fail(BALLENV_EINVAL, "patch geometry needs %zu bytes of shared memory", smem);
  if (h->cfg.precision == BALLENV_F64) {
    CUDA_TRY(cudaFuncSetAttribute(ballenv_patch_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ballenv_patch_kernel<double><<<(unsigned)h->n, kPatchThreads, smem, s>>>(h->base, tb, out, out_format == BALLENV_OBS_U8);
  } else {
    CUDA_TRY(cudaFuncSetAttribute(ballenv_patch_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ballenv_patch_kernel<float><<<(unsigned)h->n, kPatchThreads, smem, s>>>(h->base, tb, out, out_format == BALLENV_OBS_U8);
  }
  h->launches += 1;
  CUDA_TRY(cudaGetLastError());
  return BALLENV_OK;
}

int ballenv_step(BallenvHandle* h, const void* actions, int action_kind, void* obs_out, void* reward_out,
                 uint8_t* done_out, ballenv_stream_t stream) {
  if (h == nullptr || actions == nullptr) return fail(BALLENV_EINVAL, "NULL argument");
  if (action_bytes(action_kind) == 0) return fail(BALLENV_EINVAL, "unknown action_kind %d", action_kind);
  DeviceGuard guard(h->device);
  Params p = h->base;
  p.mode = kModeStep;
  p.n_steps = 1;
  p.obs_all_steps = 0;
  p.actions = actions;
  p.action_kind = action_kind;
  p.obs = obs_out;
  p.reward = reward_out;
  p.done = done_out;
  p.reset_tape = h->reset_tape;
  if (action_kind == BALLENV_ACT_XY_F32 || action_kind == BALLENV_ACT_XY_F64)   // raw (dx, dy): the agent may leave the integers
    CUDA_TRY(cudaMemsetAsync(const_cast<uint32_t*>(h->base.state_dirty), 1, 4, (cudaStream_t)stream));
  if (h->step_tape != nullptr) {
    if (h->step_tape_pos >= h->step_tape_steps)
      return fail(BALLENV_ESTATE, "step tape exhausted after %lld steps", h->step_tape_steps);
    p.step_tape = h->step_tape + (size_t)h->step_tape_pos * (size_t)h->n * h->cfg.dynamic_obstacles * 2;
    h->step_tape_pos += 1;
  }
  return launch(h, p, (cudaStream_t)stream);
}

int ballenv_rollout_policy(BallenvHandle* h, const BallenvPolicyMLP* pol, int32_t n_steps, const float* first_obs,
                           float* obs_out, int64_t* actions_out, void* reward_out, uint8_t* done_out, float* policy_out,
                           ballenv_stream_t stream) {
  if (h == nullptr || pol == nullptr || first_obs == nullptr || obs_out == nullptr || actions_out == nullptr)
    return fail(BALLENV_EINVAL, "NULL argument");
  if (pol->fc1_weight == nullptr || pol->fc1_bias == nullptr || pol->action_weight == nullptr || pol->action_bias == nullptr)
    return fail(BALLENV_EINVAL, "NULL policy parameter");
  if (n_steps < 0) return fail(BALLENV_EINVAL, "n_steps < 0");
  if (n_steps == 0) return BALLENV_OK;
  const int n_in = obs_row_elems(h->cfg);
  if (pol->n_inputs != n_in) return fail(BALLENV_EINVAL, "the policy takes %d inputs, an observation row has %d", pol->n_inputs, n_in);
  if (pol->hidden < 8 || pol->hidden > 512 || (pol->hidden & 7) != 0) return fail(BALLENV_EINVAL, "hidden must be a multiple of 8 in 8..512");
  DeviceGuard guard(h->device);
  Params p = h->base;
  p.mode = kModeStep;
  p.n_steps = n_steps;
  p.obs_all_steps = 1;
  p.actions = nullptr;
  p.action_kind = BALLENV_ACT_INDEX_I64;
  p.obs = obs_out;
  p.reward = reward_out;
  p.done = done_out;
  p.reset_tape = h->reset_tape;
  p.step_tape = h->step_tape;
  p.pol_fc1_w = pol->fc1_weight;
  p.pol_fc1_b = pol->fc1_bias;
  p.pol_act_w = pol->action_weight;
  p.pol_act_b = pol->action_bias;
  p.pol_hidden = pol->hidden;
  p.pol_greedy = pol->greedy ? 1 : 0;
  p.pol_val_w = pol->value_weight;
  p.pol_val_b = pol->value_bias;
  if (policy_out != nullptr && (pol->value_weight == nullptr || pol->value_bias == nullptr))
    return fail(BALLENV_EINVAL, "policy_out needs the value head's weight and bias");
  LeanLauncher fn = policy_launcher(h, p);
  if (fn == nullptr)
    return fail(BALLENV_ESTATE, "no policy-in-the-loop kernel for this configuration (production mode, float32 rows, WINDOW 5 "
                                "or 10, at most 64 obstacles with at least one moving): drive ballenv_step from the caller's "
                                "policy instead");
  if (4 * lean::policy_smem_floats(n_in, pol->hidden) > 160 * 1024) return fail(BALLENV_EINVAL, "the policy does not fit in shared memory");
  const size_t n = (size_t)h->n;
  const size_t rew_b = 4;
  // the kernel indexes the [T][n] arrays with 32 bits: at most (2^31 - 1) / n steps per launch
  const int64_t max_t = n > 0 ? (int64_t)0x7fffffff / (int64_t)n : n_steps;
  if (max_t < 1) return fail(BALLENV_EINVAL, "too many environments for one launch");
  for (int64_t t0 = 0; t0 < n_steps; t0 += max_t) {
    const int64_t tn = n_steps - t0 < max_t ? n_steps - t0 : max_t;
    p.n_steps = (int)tn;
    p.pol_first_obs = t0 == 0 ? first_obs : obs_out + (size_t)(t0 - 1) * n * n_in;
    p.pol_actions = reinterpret_cast<long long*>(actions_out) + (size_t)t0 * n;
    p.pol_out = policy_out ? policy_out + (size_t)t0 * n * 10 : nullptr;
    p.obs = obs_out + (size_t)t0 * n * n_in;
    p.reward = reward_out ? (char*)reward_out + (size_t)t0 * n * rew_b : nullptr;
    p.done = done_out ? done_out + (size_t)t0 * n : nullptr;
    p.obs_row_bytes = (long long)n_in * 4;
    p.obs_step_bytes = (long long)n * p.obs_row_bytes;
    fn(p, (unsigned)((p.n + kLeanEnvsPerBlock - 1) / kLeanEnvsPerBlock), (cudaStream_t)stream);
    h->launches += 1;
    CUDA_TRY(cudaGetLastError());
  }
  return BALLENV_OK;
}

int ballenv_step_many(BallenvHandle* h, const void* actions, int action_kind, int32_t n_steps, void* obs_out,
                      int32_t obs_all_steps, void* reward_out, uint8_t* done_out, ballenv_stream_t stream) {
  if (h == nullptr || actions == nullptr) return fail(BALLENV_EINVAL, "NULL argument");
  const int ab = action_bytes(action_kind);
  if (ab == 0) return fail(BALLENV_EINVAL, "unknown action_kind %d", action_kind);
  if (n_steps < 0) return fail(BALLENV_EINVAL, "n_steps < 0");
  const size_t n = (size_t)h->n;
  const size_t obs_row = (size_t)obs_row_elems(h->cfg) * obs_elem_bytes(h->cfg);
  const size_t rew_b = h->cfg.precision == BALLENV_F64 ? 8 : 4;
  if (n_steps == 0) return BALLENV_OK;
  {
    // Open-loop rollout: one launch advances all n_steps with the state held on chip (ballenv_kernel's loop).
    DeviceGuard guard(h->device);
    Params p = h->base;
    p.mode = kModeStep;
    p.n_steps = n_steps;
    p.obs_all_steps = obs_all_steps ? 1 : 0;
    p.actions = actions;
    p.action_kind = action_kind;
    p.obs = obs_out;
    p.reward = reward_out;
    p.done = done_out;
    p.reset_tape = h->reset_tape;
    p.step_tape = h->step_tape;
    if (!h->no_rollout && (fast_eligible(h, p) || lean_launcher(h, p) != nullptr)) {
      // the kernel indexes the [T][n] arrays with 32 bits: at most (2^31 - 1) / n steps per launch
      const int64_t max_t = n > 0 ? (int64_t)0x7fffffff / (int64_t)n : n_steps;
      if (max_t < 1) return fail(BALLENV_EINVAL, "too many environments for one launch");
      for (int64_t t0 = 0; t0 < n_steps; t0 += max_t) {
        const int64_t tn = n_steps - t0 < max_t ? n_steps - t0 : max_t;
        p.n_steps = (int)tn;
        p.actions = (const char*)actions + (size_t)t0 * n * ab;
        p.obs = obs_all_steps ? (char*)obs_out + (size_t)t0 * n * obs_row : obs_out;
        p.reward = reward_out ? (char*)reward_out + (size_t)t0 * n * rew_b : nullptr;
        p.done = done_out ? done_out + (size_t)t0 * n : nullptr;
        int rc = launch(h, p, (cudaStream_t)stream);
        if (rc != BALLENV_OK) return rc;
      }
      return BALLENV_OK;
    }
  }
  for (int t = 0; t < n_steps; ++t) {
    void* obs_t = nullptr;
    if (obs_out != nullptr) {
      if (obs_all_steps) obs_t = (char*)obs_out + (size_t)t * n * obs_row;
      else if (t == n_steps - 1) obs_t = obs_out;
    }
    int rc = ballenv_step(h, (const char*)actions + (size_t)t * n * ab, action_kind, obs_t,
                          reward_out ? (char*)reward_out + (size_t)t * n * rew_b : nullptr,
                          done_out ? done_out + (size_t)t * n : nullptr, stream);
    if (rc != BALLENV_OK) return rc;
  }
  return BALLENV_OK;
}

int ballenv_step_host(BallenvHandle* h, const void* actions_host, int action_kind, void* obs_host, void* reward_host,
                      uint8_t* done_host, ballenv_stream_t stream) {
  if (h == nullptr || actions_host == nullptr) return fail(BALLENV_EINVAL, "NULL argument");
  const int ab = action_bytes(action_kind);
  if (ab == 0) return fail(BALLENV_EINVAL, "unknown action_kind %d", action_kind);
  DeviceGuard guard(h->device);
  int rc = ensure_stage(h, action_kind);
  if (rc != BALLENV_OK) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t n = (size_t)h->n;
  char* st = h->stage;
  CUDA_TRY(cudaMemcpyAsync(st + h->stage_act, actions_host, n * ab, cudaMemcpyHostToDevice, s));
  rc = ballenv_step(h, st + h->stage_act, action_kind, obs_host ? st + h->stage_obs : nullptr,
                    reward_host ? st + h->stage_rew : nullptr, done_host ? (uint8_t*)(st + h->stage_done) : nullptr,
                    stream);
  if (rc != BALLENV_OK) return rc;
  if (obs_host)
    CUDA_TRY(cudaMemcpyAsync(obs_host, st + h->stage_obs, n * obs_row_elems(h->cfg) * obs_elem_bytes(h->cfg),
                             cudaMemcpyDeviceToHost, s));
  if (reward_host)
    CUDA_TRY(cudaMemcpyAsync(reward_host, st + h->stage_rew, n * (h->cfg.precision == BALLENV_F64 ? 8 : 4),
                             cudaMemcpyDeviceToHost, s));
  if (done_host) CUDA_TRY(cudaMemcpyAsync(done_host, st + h->stage_done, n, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  return BALLENV_OK;
}

int ballenv_step_many_host(BallenvHandle* h, const void* actions_host, int action_kind, int32_t n_steps, void* obs_host,
                           void* reward_host, uint8_t* done_host, ballenv_stream_t stream) {
  if (h == nullptr || actions_host == nullptr) return fail(BALLENV_EINVAL, "NULL argument");
  const int ab = action_bytes(action_kind);
  if (ab == 0) return fail(BALLENV_EINVAL, "unknown action_kind %d", action_kind);
  if (n_steps < 0) return fail(BALLENV_EINVAL, "n_steps < 0");
  if (n_steps == 0) return BALLENV_OK;
  DeviceGuard guard(h->device);
  const size_t n = (size_t)h->n;
  const size_t act_b = n * (size_t)ab;
  const size_t obs_b = n * (size_t)obs_row_elems(h->cfg) * obs_elem_bytes(h->cfg);
  const size_t rew_b = n * (h->cfg.precision == BALLENV_F64 ? 8 : 4);
  const size_t done_b = n;
  // The actions of all steps are known up front: the rollout proceeds in chunks of `tc` steps - one copy in, ONE launch
  // of the rollout kernel (ballenv_step_many), three copies out per chunk - with two device staging sets, so that the
  // copies out of chunk c overlap the kernel of chunk c + 1 and the copy in of chunk c + 2.  Every step's actions and
  // every step's rows / rewards / dones cross the bus; the state never leaves the device.
  const size_t per_step = align_up(act_b, 256) + align_up(obs_b, 256) + align_up(rew_b, 256) + align_up(done_b, 256);
  int tc = (int)((size_t)(192u << 20) / per_step);   // staging budget per set
  tc = tc < 1 ? 1 : (tc > 64 ? 64 : tc);
  if (tc > (n_steps + 5) / 6) tc = (n_steps + 5) / 6;   // at least six chunks per call: something to overlap
  const size_t o_act = 0, o_obs = align_up(act_b * tc, 256), o_rew = o_obs + align_up(obs_b * tc, 256),
               o_done = o_rew + align_up(rew_b * tc, 256), set_b = o_done + align_up(done_b * tc, 256);
  if (h->many_bytes < 2 * set_b) {
    if (h->many) cudaFree(h->many);
    h->many = nullptr;
    h->many_bytes = 0;
    CUDA_TRY(cudaMalloc(&h->many, 2 * set_b));
    h->many_bytes = 2 * set_b;
  }
  if (h->s_in == nullptr) {
    CUDA_TRY(cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      CUDA_TRY(cudaEventCreateWithFlags(&h->ev_in[i], cudaEventDisableTiming));
      CUDA_TRY(cudaEventCreateWithFlags(&h->ev_k[i], cudaEventDisableTiming));
      CUDA_TRY(cudaEventCreateWithFlags(&h->ev_out[i], cudaEventDisableTiming));
    }
    CUDA_TRY(cudaEventCreateWithFlags(&h->ev_start, cudaEventDisableTiming));
  }
  cudaStream_t s = (cudaStream_t)stream;
  CUDA_TRY(cudaEventRecord(h->ev_start, s));          // the copy streams start behind whatever the caller had enqueued
  CUDA_TRY(cudaStreamWaitEvent(h->s_in, h->ev_start, 0));
  CUDA_TRY(cudaStreamWaitEvent(h->s_out, h->ev_start, 0));
  int c = 0;
  for (int t0 = 0; t0 < n_steps; t0 += tc, ++c) {
    const int tn = n_steps - t0 < tc ? n_steps - t0 : tc;
    const int b = c & 1;
    char* st = h->many + (size_t)b * set_b;
    if (c >= 2) CUDA_TRY(cudaStreamWaitEvent(h->s_in, h->ev_k[b], 0));      // the kernel of chunk c - 2 has read this set's actions
    CUDA_TRY(cudaMemcpyAsync(st + o_act, (const char*)actions_host + (size_t)t0 * act_b, act_b * tn, cudaMemcpyHostToDevice,
                             h->s_in));
    CUDA_TRY(cudaEventRecord(h->ev_in[b], h->s_in));
    CUDA_TRY(cudaStreamWaitEvent(s, h->ev_in[b], 0));
    if (c >= 2) CUDA_TRY(cudaStreamWaitEvent(s, h->ev_out[b], 0));          // the outputs of chunk c - 2 have left this set
    int rc = ballenv_step_many(h, st + o_act, action_kind, tn, obs_host ? st + o_obs : nullptr, 1,
                               reward_host ? st + o_rew : nullptr, done_host ? (uint8_t*)(st + o_done) : nullptr, stream);
    if (rc != BALLENV_OK) return rc;
    CUDA_TRY(cudaEventRecord(h->ev_k[b], s));
    CUDA_TRY(cudaStreamWaitEvent(h->s_out, h->ev_k[b], 0));
    if (obs_host)
      CUDA_TRY(cudaMemcpyAsync((char*)obs_host + (size_t)t0 * obs_b, st + o_obs, obs_b * tn, cudaMemcpyDeviceToHost, h->s_out));
    if (reward_host)
      CUDA_TRY(cudaMemcpyAsync((char*)reward_host + (size_t)t0 * rew_b, st + o_rew, rew_b * tn, cudaMemcpyDeviceToHost,
                               h->s_out));
    if (done_host)
      CUDA_TRY(cudaMemcpyAsync(done_host + (size_t)t0 * done_b, st + o_done, done_b * tn, cudaMemcpyDeviceToHost, h->s_out));
    CUDA_TRY(cudaEventRecord(h->ev_out[b], h->s_out));
  }
  CUDA_TRY(cudaStreamSynchronize(h->s_out));
  CUDA_TRY(cudaStreamSynchronize(s));
  return BALLENV_OK;
}

int ballenv_set_draw_tape(BallenvHandle* h, const uint32_t* step_tape, int64_t n_steps, const uint32_t* reset_tape,
                          int64_t n_episodes, int32_t attempts) {
  if (h == nullptr) return fail(BALLENV_EINVAL, "handle is NULL");
  DeviceGuard guard(h->device);
  CUDA_TRY(cudaDeviceSynchronize());
  if (h->step_tape) cudaFree(h->step_tape);
  if (h->reset_tape) cudaFree(h->reset_tape);
  h->step_tape = nullptr;
  h->reset_tape = nullptr;
  h->step_tape_steps = h->step_tape_pos = 0;
  h->base.reset_tape_episodes = 0;
  h->base.tape_attempts = 0;
  h->base.reset_tape_width = 0;
  if (step_tape != nullptr && n_steps > 0 && h->cfg.dynamic_obstacles > 0) {
    const size_t bytes = (size_t)n_steps * (size_t)h->n * h->cfg.dynamic_obstacles * 2 * sizeof(uint32_t);
    CUDA_TRY(cudaMalloc(&h->step_tape, bytes));
    CUDA_TRY(cudaMemcpy(h->step_tape, step_tape, bytes, cudaMemcpyHostToDevice));
    h->step_tape_steps = n_steps;
  }
  if (reset_tape != nullptr && n_episodes > 0) {
    if (h->cfg.ruleset != BALLENV_RULESET_GYM) return fail(BALLENV_EINVAL, "reset tapes exist for the gym ruleset only");
    if (attempts < 1) return fail(BALLENV_EINVAL, "attempts must be >= 1");
    const int width = 4 + 2 * attempts * h->cfg.static_obstacles + 2 * h->cfg.dynamic_obstacles;
    const size_t bytes = (size_t)n_episodes * (size_t)h->n * width * sizeof(uint32_t);
    CUDA_TRY(cudaMalloc(&h->reset_tape, bytes));
    CUDA_TRY(cudaMemcpy(h->reset_tape, reset_tape, bytes, cudaMemcpyHostToDevice));
    h->base.reset_tape_episodes = n_episodes;
    h->base.tape_attempts = attempts;
    h->base.reset_tape_width = width;
  }
  return BALLENV_OK;
}

int ballenv_stats(BallenvHandle* h, double* out, ballenv_stream_t stream) {
  if (h == nullptr || out == nullptr) return fail(BALLENV_EINVAL, "NULL argument");
  DeviceGuard guard(h->device);
  CUDA_TRY(cudaMemcpyAsync(out, h->base.stats, sizeof(double) * BALLENV_NUM_STATS, cudaMemcpyDeviceToHost,
                           (cudaStream_t)stream));
  CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
  return BALLENV_OK;
}

int ballenv_stats_reset(BallenvHandle* h, ballenv_stream_t stream) {
  if (h == nullptr) return fail(BALLENV_EINVAL, "handle is NULL");
  DeviceGuard guard(h->device);
  CUDA_TRY(cudaMemsetAsync(h->base.stats, 0, sizeof(double) * BALLENV_NUM_STATS, (cudaStream_t)stream));
  return BALLENV_OK;
}

int ballenv_error_flags(BallenvHandle* h, uint32_t* out, ballenv_stream_t stream) {
  if (h == nullptr || out == nullptr) return fail(BALLENV_EINVAL, "NULL argument");
  DeviceGuard guard(h->device);
  CUDA_TRY(cudaMemcpyAsync(out, h->base.errors, sizeof(uint32_t), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  CUDA_TRY(cudaMemsetAsync(h->base.errors, 0, sizeof(uint32_t), (cudaStream_t)stream));
  CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
  return BALLENV_OK;
}

int64_t ballenv_launch_count(BallenvHandle* h) { return h ? h->launches : 0; }

int ballenv_kernel_variant(BallenvHandle* h, int action_kind, int32_t n_steps) {
  if (h == nullptr) return fail(BALLENV_EINVAL, "handle is NULL");
  Params p = h->base;
  p.mode = kModeStep;
  p.n_steps = n_steps > 1 ? n_steps : 1;
  p.action_kind = action_kind;
  p.obs = h->arena;   // any non-null pointer: rows requested
  p.reset_tape = h->reset_tape;
  p.step_tape = h->step_tape;
  if (p.n_steps > 1 && h->no_rollout) p.n_steps = 1;
  int lanes = 1;
  if (lean_launcher(h, p, &lanes) != nullptr) return lanes == 2 ? BALLENV_KERNEL_LEAN2 : BALLENV_KERNEL_LEAN;
  return fast_eligible(h, p) ? BALLENV_KERNEL_ROLES : BALLENV_KERNEL_GENERIC;
}

namespace {
// R_t = r_t + gamma * R_{t+1} * (1 - done_t) from the end of a [T][n] slice, one thread per environment: the loop of
// examples/ball_cnn_ac3.py:228-230 for N trajectories at once.  fp32, rounded like the tensor expression
// reward[t] + gamma * R * mask (no contraction).
__global__ void __launch_bounds__(256) discounted_returns_kernel(const float* __restrict__ reward, const uint8_t* __restrict__ done,
                                                                 const float* __restrict__ bootstrap, float gamma, int T,
                                                                 long long n, float* __restrict__ out) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  float R = bootstrap != nullptr ? bootstrap[e] : 0.0f;
  for (int t = T - 1; t >= 0; --t) {
    const size_t i = (size_t)t * (size_t)n + (size_t)e;
    R = __fadd_rn(reward[i], __fmul_rn(__fmul_rn(gamma, R), done[i] ? 0.0f : 1.0f));
    out[i] = R;
  }
}
}  // namespace

int ballenv_discounted_returns(const float* reward, const uint8_t* done, const float* bootstrap, float gamma, int32_t n_steps,
                               int64_t n, float* out, ballenv_stream_t stream) {
  if (reward == nullptr || done == nullptr || out == nullptr) return fail(BALLENV_EINVAL, "NULL argument");
  if (n_steps < 0 || n < 0) return fail(BALLENV_EINVAL, "negative size");
  if (n_steps == 0 || n == 0) return BALLENV_OK;
  discounted_returns_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(reward, done, bootstrap, gamma, n_steps, n, out);
  CUDA_TRY(cudaGetLastError());
  return BALLENV_OK;
}

namespace {
int a2c_grid(int n_in, int hidden, int tile, int nw, int64_t n_samples, size_t* smem_out) {
  const size_t smem = a2c::smem_bytes(n_in, hidden, tile, nw);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int per_sm = (int)((220 * 1024) / (smem + 1024));
  if (per_sm > 5) per_sm = 5;
  if (per_sm < 1) per_sm = 1;
  int64_t grid = (int64_t)sms * per_sm;
  const int64_t tiles = (n_samples + tile - 1) / tile;
  if (grid > tiles) grid = tiles;
  if (smem_out != nullptr) *smem_out = smem;
  return (int)(grid < 1 ? 1 : grid);
}
int a2c_tile(int hidden) { return hidden <= 128 ? 128 : (hidden + 31) / 32 * 32; }
}  // namespace

int64_t ballenv_a2c_workspace_bytes(int32_t n_inputs, int32_t hidden, int64_t n_samples) {
  if (n_inputs < 1 || hidden < 4 || hidden > a2c::kMaxHidden || (hidden & 3) != 0 || n_samples < 0) return fail(BALLENV_EINVAL, "bad policy shape");
  const int grid = a2c_grid(n_inputs, hidden, a2c_tile(hidden), (n_inputs + 31) / 32, n_samples > 0 ? n_samples : 1, nullptr);
  return (int64_t)grid * (a2c::n_params(n_inputs, hidden) + 1) * 4;
}

int ballenv_a2c_grads(const BallenvA2CUpdate* u, const float* obs, const int64_t* actions, const float* returns,
                      int64_t n_samples, void* workspace, int64_t workspace_bytes, ballenv_stream_t stream) {
  if (u == nullptr || obs == nullptr || actions == nullptr || returns == nullptr || workspace == nullptr)
    return fail(BALLENV_EINVAL, "NULL argument");
  const void* ptrs[] = {u->fc1_weight, u->fc1_bias, u->action_weight, u->action_bias, u->value_weight, u->value_bias,
                        u->fc1_weight_grad, u->fc1_bias_grad, u->action_weight_grad, u->action_bias_grad, u->value_weight_grad,
                        u->value_bias_grad, u->loss};
  for (const void* q : ptrs)
    if (q == nullptr) return fail(BALLENV_EINVAL, "NULL policy parameter / gradient pointer");
  const int nw = (u->n_inputs + 31) / 32;
  if (u->n_inputs < 1 || (nw != 1 && nw != 4)) return fail(BALLENV_EINVAL, "n_inputs must be 4 + W*W for WINDOW = 5 (29) or up to 128 (WINDOW = 10: 104)");
  if (u->hidden < 4 || u->hidden > a2c::kMaxHidden || (u->hidden & 3) != 0) return fail(BALLENV_EINVAL, "hidden must be a multiple of 4 in 4..%d", a2c::kMaxHidden);
  if (n_samples < 1) return fail(BALLENV_EINVAL, "n_samples < 1");
  const int tile = a2c_tile(u->hidden);
  size_t smem = 0;
  const int grid = a2c_grid(u->n_inputs, u->hidden, tile, nw, n_samples, &smem);
  const int P = a2c::n_params(u->n_inputs, u->hidden);
  if (workspace_bytes < (int64_t)grid * (P + 1) * 4) return fail(BALLENV_EINVAL, "workspace too small: ballenv_a2c_workspace_bytes");
  if (smem > 220 * 1024) return fail(BALLENV_EINVAL, "the policy does not fit in shared memory");
  a2c::Args a;
  a.n_in = u->n_inputs;
  a.hidden = u->hidden;
  a.tile = tile;
  a.n_samples = n_samples;
  a.fc1_w = u->fc1_weight;
  a.fc1_b = u->fc1_bias;
  a.act_w = u->action_weight;
  a.act_b = u->action_bias;
  a.val_w = u->value_weight;
  a.val_b = u->value_bias;
  a.obs = obs;
  a.action = reinterpret_cast<const long long*>(actions);
  a.ret = returns;
  a.ret_stats = u->returns_stats;
  a.pv = u->policy_out;
  a.partial = reinterpret_cast<float*>(workspace);
  cudaStream_t s = (cudaStream_t)stream;
  if (nw == 1) {
    CUDA_TRY(cudaFuncSetAttribute(a2c::grad_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    a2c::grad_kernel<1><<<grid, tile, smem, s>>>(a);
  } else {
    CUDA_TRY(cudaFuncSetAttribute(a2c::grad_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    a2c::grad_kernel<4><<<grid, tile, smem, s>>>(a);
  }
  CUDA_TRY(cudaGetLastError());
  a2c::Outs o;
  o.fc1_w = u->fc1_weight_grad;
  o.fc1_b = u->fc1_bias_grad;
  o.act_w = u->action_weight_grad;
  o.act_b = u->action_bias_grad;
  o.val_w = u->value_weight_grad;
  o.val_b = u->value_bias_grad;
  o.loss = u->loss;
  a2c::reduce_kernel<<<(P + 1 + 31) / 32, 256, 0, s>>>(a.partial, grid, u->n_inputs, u->hidden, o);
  CUDA_TRY(cudaGetLastError());
  return BALLENV_OK;
}

int ballenv_selftest(int which, int64_t arg, int device, int64_t* mismatches) {
  if (mismatches == nullptr) return fail(BALLENV_EINVAL, "mismatches is null");
  if (which != 0 && which != 1) return fail(BALLENV_EINVAL, "unknown self-test %d", which);
  if (which == 0 && (arg < 0 || arg > (1ll << 22))) return fail(BALLENV_EINVAL, "sqrt_int22 is defined for 0 <= s < 2^22");
  if (which == 1 && (arg < 0 || arg > (1ll << 34))) return fail(BALLENV_EINVAL, "at most 2^34 pairs");
  DeviceGuard guard(device);
  unsigned long long* d_bad = nullptr;
  CUDA_TRY(cudaMalloc(&d_bad, sizeof(*d_bad)));
  cudaMemset(d_bad, 0, sizeof(*d_bad));
  if (which == 0) selftest_sqrt_kernel<<<(unsigned)((arg + 255) / 256), 256>>>(arg, d_bad);
  else selftest_div_kernel<<<(unsigned)((arg + 255) / 256), 256>>>(arg, d_bad);
  unsigned long long bad = 0;
  cudaError_t e = cudaMemcpy(&bad, d_bad, sizeof(bad), cudaMemcpyDeviceToHost);
  cudaFree(d_bad);
  if (e != cudaSuccess) return fail(BALLENV_ECUDA, "self-test failed: %s", cudaGetErrorString(e));
  *mismatches = (int64_t)bad;
  return BALLENV_OK;
}

}  // extern "C"
