// Fused step + window-observe kernel of the batched ball environment (sm_100a).
//
// One launch advances every environment of a handle by one step (ballenv_step) or by T steps with the state held
// on chip (ballenv_step_many, the rollout loop):
//   agent move + wall clamp            gym_ballenv/envs/ballenv_env.py:236-259 | ballenv_pygame.py:652-665
//   obstacle motion                    ballenv_env.py:262-264, 323-353
//   distance, goal / obstacle tests,   ballenv_env.py:268-286, 200-229, 179-191 | ballenv_pygame.py:668-706
//   reward, accumulated reward, done
//   TimeLimit(1000) truncation         gym_ballenv/__init__.py:7 (gym 0.10.9 wrapper, restated)
//   auto-reset of finished envs        ballenv_env.py:113-167 | ballenv_pygame.py:460-513 (Philox draws)
//   WINDOW x WINDOW occupancy + goal   examples/ball_cnn_ac3.py:330-352, 384-412 (incl. the row-offset quirk :409)
//   quadrant observation
//
// Mapping (B200: 148 SMs want several hundred thousand threads in flight; 64 K environments alone are too few):
//   * a block of 288 threads owns 32 consecutive environments.  Warp 0 is the scalar warp: thread t owns
//     environment t's scalars (agent, goal, distance, reward, counters), struct-of-arrays, so every load and
//     store of the warp is one full 128-byte line.  Warps 1..8 are obstacle threads.
//   * obstacles live env-major, [n][K padded to 4].  The block's quads form one linear slot space (static quads
//     first, then dynamic ones, exactly as they lie in memory); thread s owns slot s and moves it with one 128-bit
//     load and one 128-bit store per field (x, y, meta), so consecutive threads issue consecutive 128-bit accesses
//     and a warp is purely static or purely dynamic.  One Philox4x32-10 block feeds the four obstacles of a quad.
//     Draws and moves do not depend on the agent and overlap the scalar warp's action -> move -> clamp chain.
//   * most obstacles are far from the agent.  Each thread bounding-box tests its quad (one branch for the four
//     tests); the rare near obstacles are hit-tested and appended to a shared-memory list, and the obstacle threads
//     then rasterise the list one (obstacle, window row) item per thread with exactly the reference's arithmetic
//     (dx*dx + dy*dy against the radius sum), OR-ing row masks into the block's observation bit-stream in shared
//     memory.  The grid is therefore bit-exact by construction and the work is balanced wherever obstacles cluster.
//     Meanwhile the scalar warp computes distance, reward, flags, done and the episode statistics in fp64.
//   * the block's observation span [32 envs][4 + W*W] is contiguous and 128-byte aligned in global memory and its
//     element f is bit f of the bit-stream: a 16-entry float4 table expands one nibble into one 128-bit streaming
//     store.  In the rollout loop the static-quad threads (no draws, no moves) do the stores.
//   * environments that finish are reset in the same launch (reset_stage, out of line: rejection sampling per
//     obstacle, quads in parallel, second list raster); the hot loop is left for it and re-entered after a reload.
// Nothing here is a dense contraction: no tensor cores.
#pragma once
#include <stdint.h>

#include "../../include/ballenv.h"
#include "ballenv_rng.cuh"

namespace ballenv {

constexpr int kLanes = 8;                          // lanes per environment
constexpr int kEnvsPerBlock = 32;                  // = one warp in the thread-per-env scalar phase
constexpr int kLaneThreads = kLanes * kEnvsPerBlock;   // 256 obstacle threads (warps 1..8)
constexpr int kBlock = 32 + kLaneThreads;          // + warp 0, the scalar warp: 288 threads
#ifndef BALLENV_MINBLOCKS
#define BALLENV_MINBLOCKS 4
#endif
// resident blocks per SM the register budget is held to; kLW = obstacle warps of the block (8, or 6 for
// configurations of at most six quads per environment: 224 threads, five blocks at the same 56 registers)
template <typename T, int kLW = kLanes>
constexpr int kMinBlocks = sizeof(T) == 4 ? (kLW == 6 ? 5 : kLW == 4 ? 7 : BALLENV_MINBLOCKS) : 2;
// near-obstacle list entries per block (typically one or two are in use; a longer list only takes L1 away from the
// spilled loop state: 768 -> 192 measured 1 % faster for C3); overflow is rasterised in-lane
constexpr int kListCap = 192;
constexpr int kMaxResetAttempts = 4096;            // the reference would loop forever on an unsatisfiable layout
constexpr int kNoHit = 0x7fffffff;

enum Mode : int { kModeStep = 0, kModeReset = 1, kModeObserve = 2 };

struct DevConfig {
  int ruleset, window, ks, kd, n_goals, goals_distinct, change_step, rd_th;
  int max_steps, auto_reset, obs_format, obs_row_elems;
  double static_penalty, dynamic_penalty;
  double world_w, world_h, radius_sum, goal_threshold, step_x, step_y;
  double reset_agent_thresh, reset_goal_thresh;  // pygame reset clearances (ballenv_pygame.py:494)
  double margin;   // bounding-box half-size beyond which an obstacle cannot touch the window or the agent
  double speed[BALLENV_MAX_DYNAMIC];
  double goal_x[BALLENV_MAX_GOALS], goal_y[BALLENV_MAX_GOALS];
  // fp32 copies so that the production instantiation never converts in the hot loop
  float f_world_w, f_world_h, f_step_x, f_step_y, f_margin;
  uint32_t rcp_qs, rcp_qd;   // ceil(2^32 / quads per environment) (0 when there is one quad: identity)
  float f_speed[BALLENV_MAX_DYNAMIC];
  float2 f_goal[BALLENV_MAX_GOALS];
  int lean_integral_speeds;  // every obstacle speed is an integer (integral coordinates then stay integral: ballenv_lean.cuh)
  uint32_t lean_cs4;         // change_step in every byte (when it fits one)
};

template <typename T>
struct CfgV;
template <>
struct CfgV<float> {
  static __device__ __forceinline__ float world_w(const DevConfig& c) { return c.f_world_w; }
  static __device__ __forceinline__ float world_h(const DevConfig& c) { return c.f_world_h; }
  static __device__ __forceinline__ float step_x(const DevConfig& c) { return c.f_step_x; }
  static __device__ __forceinline__ float step_y(const DevConfig& c) { return c.f_step_y; }
  static __device__ __forceinline__ float margin(const DevConfig& c) { return c.f_margin; }
  static __device__ __forceinline__ float speed(const DevConfig& c, int j) { return c.f_speed[j]; }
  static __device__ __forceinline__ float2 goal(const DevConfig& c, int i) { return c.f_goal[i]; }
};
template <>
struct CfgV<double> {
  static __device__ __forceinline__ double world_w(const DevConfig& c) { return c.world_w; }
  static __device__ __forceinline__ double world_h(const DevConfig& c) { return c.world_h; }
  static __device__ __forceinline__ double step_x(const DevConfig& c) { return c.step_x; }
  static __device__ __forceinline__ double step_y(const DevConfig& c) { return c.step_y; }
  static __device__ __forceinline__ double margin(const DevConfig& c) { return c.margin; }
  static __device__ __forceinline__ double speed(const DevConfig& c, int j) { return c.speed[j]; }
  static __device__ __forceinline__ double2 goal(const DevConfig& c, int i) { return make_double2(c.goal_x[i], c.goal_y[i]); }
};

struct Params {
  DevConfig cfg;
  long long n, stride;           // environments, element stride of the per-env scalar arrays
  long long stat_stride, dyn_stride;  // elements per environment row of the obstacle arrays (K padded to 4)
  uint32_t g0, k0, k1;
  int mode, action_kind;
  int n_steps;        // steps advanced by this launch (rollout loop; > 1 only with the fast specialisation)
  int obs_all_steps;  // rollout: observation rows of every step ([T][n][row]) or only of the last one ([n][row])
  // BALLENV_DEBUG_SKIP (profiling experiments, tools/skip_experiment.py): 1 exit, 2 no moves, 4 no near tests,
  // 8 no store, 16 no fp64.  Compiling the tests of this word out was measured SLOWER (tools/ab.sh: 10.7 -> 11.4 us
  // per step for C3 - a different ptxas schedule), so they stay in; the word is 0 unless the variable is set.
  int debug;
  // derived per launch by the host (ballenv_capi.cu: launch) so that the step loop re-reads one constant instead of
  // recomputing it when registers are short
  long long obs_row_bytes, obs_step_bytes;   // bytes of one observation row; of one step's rows ([T][n][row]) or 0
  int n_stat, n_slot;                        // static-quad threads / all busy obstacle threads of a block (32 environments)
  int sq;                                    // static quads per static-quad thread: 1, or 2 (kernels instantiated with kSQ = 2)
  void *agent_x, *agent_y, *goal_x, *goal_y;
  double *dist, *total, *acc;
  int *ep_len;
  uint32_t *episode, *tick;
  void *stat_x, *stat_y, *dyn_x, *dyn_y;
  uint32_t* dyn_meta;
  uint8_t* flags;
  double* stats;
  uint32_t* errors;
  const void* actions;
  void* obs;
  void* reward;
  uint8_t* done;
  const uint8_t* reset_mask;
  const uint32_t* step_tape;   // [n][kd][2] words of this step, or null
  const uint32_t* reset_tape;  // [episodes][n][width], or null
  long long reset_tape_episodes;
  int tape_attempts, reset_tape_width;
  // thread-per-environment kernels (ballenv_lean.cuh): Philox round keys k + r * W of the handle's seed, and the table
  // of window column masks (LeanTab<W>, in the arena)
  uint32_t rk[20];
  const uint16_t* lean_tab;
  // 0: every coordinate of the handle is known to be integral and inside the exact ranges (the lean kernels then skip
  // their per-launch test); non-zero: unknown.  Maintained by the host (ballenv_capi.cu: full resets clear it, float
  // actions and ballenv_state_written's validation set it)
  const uint32_t* state_dirty;
  // policy-in-the-loop rollouts (ballenv_rollout_policy; lean kernels instantiated with kPolicy): the MLP of
  // examples/ball_cnn_ac3.py:109-146 in nn.Linear layout, the observation the first step acts on, the actions taken
  const float *pol_fc1_w, *pol_fc1_b;   // [hidden][4 + W*W], [hidden]
  const float *pol_act_w, *pol_act_b;   // [9][hidden], [9]
  const float* pol_first_obs;           // float32 [n][4 + W*W]
  long long* pol_actions;               // int64 [T][n]
  const float *pol_val_w, *pol_val_b;   // value head [1][hidden], [1] (only for pol_out), or nullptr
  float* pol_out;                       // float32 [T][n][10]: the 9 action probabilities and the value each step acted on, or nullptr
  int pol_hidden, pol_greedy;
};

// ---- arithmetic that must not be contracted into FMAs (the reference squares, then adds) -------------------
__device__ __forceinline__ float r_mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float r_add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float r_sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double r_mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double r_add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double r_sub(double a, double b) { return __dsub_rn(a, b); }

__device__ __forceinline__ float r_abs(float a) { return fabsf(a); }
__device__ __forceinline__ double r_abs(double a) { return fabs(a); }

// num / den, bit for bit.  A zero numerator is common (the agent did not change its distance to the goal: action
// (0, 0), or pinned at a wall) and sends the lane - and with it the scalar warp, the longest dependency chain of a
// step - through the slow path of the fp64 division; (+-0) / den is +-0 for den > 0, so those lanes divide 1 instead.
__device__ __forceinline__ double div64(double num, double den) {
  const bool zero = num == 0.0 && den > 0.0;
  double safe = zero ? 1.0 : num;
  asm("" : "+d"(safe));   // opaque: otherwise the compiler divides num itself again (its quotient is unused when zero)
  const double q = safe / den;
  return zero ? num : q;
}

__device__ __forceinline__ double dist64(double ax, double ay, double bx, double by) {
  const double dx = __dsub_rn(ax, bx), dy = __dsub_rn(ay, by);
  return sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
}

// sqrt((double)s) for an integer-valued float 0 <= s < 2^22, bit for bit (ballenv_selftest(0, ...) compares every
// s of that range with sqrt()): fp32 reciprocal-square-root seed, one coupled Newton step in fp64 for g ~ sqrt(s)
// and h ~ 1 / (2 sqrt(s)), then g + (s - g*g) * h with the residual as a single fused operation - the value before
// the final rounding is within 2^-80 of sqrt(s), far inside the distance of any such root from a rounding
// boundary.  Four dependent fp64 operations instead of the library routine's dozen - and the distance to the
// goal sits on the scalar warp's chain, the longest of a step.
__device__ __forceinline__ double sqrt_int22(float s) {
  float rs;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rs) : "f"(s));
  const double sd = (double)s, h0 = 0.5 * (double)rs, g0 = sd * (double)rs;   // exact products: one shared seed error
  const double r0 = fma(-g0, h0, 0.5);
  const double g1 = fma(g0, r0, g0), h1 = fma(h0, r0, h0);
  const double g2 = fma(fma(-g1, g1, sd), h1, g1);
  return s == 0.0f ? 0.0 : g2;
}
// true if v is an integer of magnitude <= 1024: differences of two such values are exact in fp32 and the sum of
// two squared differences stays below 2^22 + 1
__device__ __forceinline__ bool small_int(float v) { return v == truncf(v) && fabsf(v) <= 1024.0f; }
__device__ __forceinline__ bool small_int(double) { return false; }

// check_overlap (ballenv_env.py:185-191): NOT (sqrt(dx^2 + dy^2) > r).  In fp32 the squares of the integral
// coordinates the env produces are exact (< 2^24), so "<= r^2" is the same predicate without the sqrt.
template <typename T>
struct Overlap;
template <>
struct Overlap<float> {
  float r2;
  __device__ explicit Overlap(double r) : r2((float)(r * r)) {}
  __device__ __forceinline__ bool operator()(float dx, float dy) const {
    return __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)) <= r2;
  }
};
template <>
struct Overlap<double> {
  double r;
  __device__ explicit Overlap(double r_) : r(r_) {}
  __device__ __forceinline__ bool operator()(double dx, double dy) const {
    return !(sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy))) > r);
  }
};

template <int W>
struct Win {
  static constexpr int kMax = W ? W : BALLENV_MAX_WINDOW;
  static constexpr int kWords = (4 + kMax * kMax + 31) / 32;   // per environment, BALLENV_OBS_BITS rows
  static constexpr int kBlockWords = 4 + kMax * kMax;         // 32 environments x (4 + W*W) bits = 4 + W*W words
  __device__ static __forceinline__ int w(int rt) { return W ? W : rt; }
};

// ---- 128-bit quad access -------------------------------------------------------------------------------------
__device__ __forceinline__ void load4(const float* p, float (&v)[4]) {
  const float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void load4(const double* p, double (&v)[4]) {
  const double2 a = reinterpret_cast<const double2*>(p)[0], b = reinterpret_cast<const double2*>(p)[1];
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
__device__ __forceinline__ void load4(const uint32_t* p, uint32_t (&v)[4]) {
  const uint4 t = *reinterpret_cast<const uint4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void store4(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void store4(double* p, const double (&v)[4]) {
  reinterpret_cast<double2*>(p)[0] = make_double2(v[0], v[1]);
  reinterpret_cast<double2*>(p)[1] = make_double2(v[2], v[3]);
}
__device__ __forceinline__ void store4(uint32_t* p, const uint32_t (&v)[4]) {
  *reinterpret_cast<uint4*>(p) = make_uint4(v[0], v[1], v[2], v[3]);
}

// ---- window raster (examples/ball_cnn_ac3.py:396-409) ----------------------------------------------------------
// Column c samples x = start_x + step_x * c.  Output row 0 and row 1 both sample y index 0 (start_y); row r >= 1
// samples start_y + step_y * (r - 1), because the reference advances cur_y after the column loop with the
// current r.  One "item" is (obstacle, y index yi): its column mask goes to row yi + 1 and, for yi = 0, row 0.
template <typename T, int W>
__device__ __forceinline__ uint32_t raster_row(T ox, T oy, T start_x, T start_y, T step_x, T step_y, int yi, int w,
                                               const Overlap<T>& ov) {
  const T dy = r_sub(r_add(start_y, r_mul(step_y, (T)yi)), oy);
  uint32_t m = 0;
  if constexpr (W == 0) {
    for (int c = 0; c < w; ++c) m |= (ov(r_sub(r_add(start_x, r_mul(step_x, (T)c)), ox), dy) ? 1u : 0u) << c;
  } else {
#pragma unroll
    for (int c = 0; c < W; ++c) m |= (ov(r_sub(r_add(start_x, r_mul(step_x, (T)c)), ox), dy) ? 1u : 0u) << c;
  }
  return m;
}

// OR a row mask of w bits into the bit-stream at bit offset off (shared memory, concurrent writers).
__device__ __forceinline__ void emit_bits(uint32_t* words, int off, int w, uint32_t m) {
  if (m == 0) return;
  const int sh = off & 31;
  atomicOr(&words[off >> 5], m << sh);
  if (sh + w > 32) {
    const uint32_t hi = m >> (32 - sh);
    if (hi) atomicOr(&words[(off >> 5) + 1], hi);
  }
}

// words is the block's bit-stream (environment el owns bits [el * nb, (el + 1) * nb)); base = el * nb.
template <typename T, int W>
__device__ __forceinline__ void raster_item(uint32_t* words, int base, T ox, T oy, T ax, T ay, T step_x, T step_y,
                                            int yi, int w, const Overlap<T>& ov) {
  const int h = w / 2;
  const T start_x = r_sub(ax, r_mul(step_x, (T)h)), start_y = r_sub(ay, r_mul(step_y, (T)h));
  const uint32_t m = raster_row<T, W>(ox, oy, start_x, start_y, step_x, step_y, yi, w, ov);
  if (yi + 1 < w) emit_bits(words, base + 4 + (yi + 1) * w, w, m);
  if (yi == 0) emit_bits(words, base + 4, w, m);
}

template <typename T>
struct Vec2;
template <>
struct Vec2<float> { typedef float2 type; };
template <>
struct Vec2<double> { typedef double2 type; };

// Shared-memory state of one block.
template <typename T, int W>
struct BlockShared {
  // observation bits of the block as ONE stream: environment el owns bits [el * nb, (el + 1) * nb), nb = 4 + W*W
  // (4 goal-quadrant bits, then the W*W cells), so output element f of the block's span is bit f.
  uint32_t words[2][Win<W>::kBlockWords];           // double-buffered by step parity (see the rollout loop)
  // agent position the window is centred on, double-buffered by step parity: the scalar warp runs ahead and
  // publishes the positions of step t + 1 while obstacle threads may still raster step t
  T ax[2][kEnvsPerBlock], ay[2][kEnvsPerBlock];
  int any_reset;                                     // scalar warp -> obstacle threads: an environment of the block resets
  int hit[kEnvsPerBlock];                            // list index of the first obstacle hit, or kNoHit
  int reset[kEnvsPerBlock];
  T near_x[kListCap], near_y[kListCap];              // near-obstacle list
  uint8_t near_env[kListCap];
  int count;
  typename Vec2<T>::type goal[BALLENV_MAX_GOALS];    // obstacle goals (args.obs_goal_position)
  typename Vec2<T>::type mv[12];                     // obstacle move table (ballenv_env.py:324)
  float4 lut[16];                                    // 4 observation bits -> 4 floats
  long long act[2][kEnvsPerBlock];                   // int64 action indices fetched one step ahead (cp.async)
  uint8_t rlist[kEnvsPerBlock];                      // reset stage: compact list of the resetting environments
  int nreset;
  // scalar-warp state parked between its two stages (keeps it out of the obstacle threads' register budget)
  T s_gx[kEnvsPerBlock], s_gy[kEnvsPerBlock], s_oax[kEnvsPerBlock], s_oay[kEnvsPerBlock];
  double s_old[kEnvsPerBlock], s_total[kEnvsPerBlock], s_acc[kEnvsPerBlock];
  int s_len[kEnvsPerBlock];
  uint32_t s_tick[kEnvsPerBlock];
};

// Rare path of the bounding-box test: hit-test the obstacle and queue it for the raster.
template <typename T, int W>
__device__ __forceinline__ void near_push(BlockShared<T, W>& sh, uint32_t* words, const DevConfig& cfg, int el, T ax, T ay,
                                       T ox, T oy, int k, bool want_hit, bool want_obs, int nb) {
  const Overlap<T> ov(cfg.radius_sum);
  if (want_hit && ov(r_sub(ax, ox), r_sub(ay, oy))) atomicMin(&sh.hit[el], k);
  if (want_obs) {
    const int idx = atomicAdd(&sh.count, 1);
    if (idx < kListCap) {
      sh.near_x[idx] = ox;
      sh.near_y[idx] = oy;
      sh.near_env[idx] = (uint8_t)el;
    } else {  // list full (obstacles piled up on the agents of this block): rasterise in-lane
      const int w = Win<W>::w(cfg.window);
      const int nyi = w > 1 ? w - 1 : 1;
      for (int yi = 0; yi < nyi; ++yi)
        raster_item<T, W>(words, el * nb, ox, oy, ax, ay, CfgV<T>::step_x(cfg), CfgV<T>::step_y(cfg), yi, w, ov);
    }
  }
}

// Bounding-box test of one obstacle against the agent (a handful of instructions; almost always false).
template <typename T, int W>
__device__ __forceinline__ void near_test(BlockShared<T, W>& sh, uint32_t* words, const DevConfig& cfg, int el, T ax,
                                          T ay, T margin, T ox, T oy, int k, bool want_hit, bool want_obs, int nb) {
  if (r_abs(r_sub(ax, ox)) <= margin && r_abs(r_sub(ay, oy)) <= margin)
    near_push<T, W>(sh, words, cfg, el, ax, ay, ox, oy, k, want_hit, want_obs, nb);
}

// Obstacle threads (lt = 0 .. kLaneThreads-1): rasterise the queued (obstacle, row) items.
template <typename T, int W>
__device__ __forceinline__ void raster_list(BlockShared<T, W>& sh, uint32_t* words, const DevConfig& cfg, int lt,
                                            int nb, int par, int nthreads = kLaneThreads) {
  const int cnt = sh.count < kListCap ? sh.count : kListCap;
  if (cnt == 0) return;
  const Overlap<T> ov(cfg.radius_sum);
  const int w = Win<W>::w(cfg.window);
  const int nyi = w > 1 ? w - 1 : 1;
  const int items = cnt * nyi;
  const T step_x = CfgV<T>::step_x(cfg), step_y = CfgV<T>::step_y(cfg);
  for (int it = lt; it < items; it += nthreads) {
    const int en = it / nyi, yi = it - en * nyi;
    const int el = sh.near_env[en];
    raster_item<T, W>(words, el * nb, sh.near_x[en], sh.near_y[en], sh.ax[par][el], sh.ay[par][el], step_x, step_y, yi, w, ov);
  }
}

// ---- draws ----------------------------------------------------------------------------------------------------
struct DrawCtx {
  const Params* p;
  long long e;       // local env index
  uint32_t g;        // global env id
  __device__ __forceinline__ uint4 reset_block(uint32_t episode, uint32_t c2) const {
    return philox4x32_10(g, episode, c2, kStreamReset, p->k0, p->k1);
  }
  // The 4 head words of a gym reset or a 2-word (x, y) pair; tape overrides Philox when it has the slot.
  __device__ __forceinline__ bool tape_row(uint32_t episode, const uint32_t*& row) const {
    if (p->reset_tape == nullptr) return false;
    if ((long long)episode >= p->reset_tape_episodes) {
      atomicOr(p->errors, (uint32_t)BALLENV_DEVERR_TAPE_EXHAUSTED);
      return false;
    }
    row = p->reset_tape + ((long long)episode * p->n + e) * p->reset_tape_width;
    return true;
  }
  __device__ __forceinline__ uint4 reset_head(uint32_t episode, uint32_t item) const {
    const uint32_t* row;
    if (item == 0 && tape_row(episode, row)) return make_uint4(row[0], row[1], row[2], row[3]);
    return reset_block(episode, (kResetHead << 28) | item);
  }
  __device__ __forceinline__ uint2 reset_static(uint32_t episode, int i, int attempt) const {
    const uint32_t* row;
    if (tape_row(episode, row)) {
      if (attempt < p->tape_attempts) {
        const int s = 4 + (i * p->tape_attempts + attempt) * 2;
        return make_uint2(row[s], row[s + 1]);
      }
      atomicOr(p->errors, (uint32_t)BALLENV_DEVERR_TAPE_EXHAUSTED);
    }
    const uint4 b = reset_block(episode, (kResetStatic << 28) | ((uint32_t)i << 16) | ((uint32_t)attempt >> 1));
    return (attempt & 1) ? make_uint2(b.z, b.w) : make_uint2(b.x, b.y);
  }
  __device__ __forceinline__ uint2 reset_dynamic(uint32_t episode, int j) const {
    const uint32_t* row;
    if (tape_row(episode, row)) {
      const int s = 4 + 2 * p->tape_attempts * p->cfg.ks + 2 * j;
      return make_uint2(row[s], row[s + 1]);
    }
    const uint4 b = reset_block(episode, (kResetDynamic << 28) | ((uint32_t)j >> 1));
    return (j & 1) ? make_uint2(b.z, b.w) : make_uint2(b.x, b.y);
  }
};

// ---- reset (ballenv_env.py:113-167 / ballenv_pygame.py:460-513) ------------------------------------------------
// Head of a reset: goal and agent position, state[2] and total_distance.  Every lane of the environment
// evaluates it (same address, same words), so no exchange is needed.
template <typename T>
struct ResetHead {
  T gx, gy, ax, ay;
  double dist, total;
};

template <typename T>
__device__ __forceinline__ ResetHead<T> reset_head_draw(const Params& p, const DrawCtx& dc, uint32_t episode) {
  const DevConfig& cfg = p.cfg;
  ResetHead<T> r;
  if (cfg.ruleset == BALLENV_RULESET_GYM) {
    const uint4 hw = dc.reset_head(episode, 0);
    r.gx = (T)__umulhi(hw.x, 500u);                                              // :115-116
    r.gy = (T)(480u + __umulhi(hw.y, 20u));
    r.ax = (T)__umulhi(hw.z, 500u);                                              // :117-118
    r.ay = (T)__umulhi(hw.w, 10u);
    // The redraw-while-closer-than-50 loop (:121-126) cannot trigger: goal_y - agent_y >= 471.
    r.dist = dist64((double)r.gx, (double)r.gy, (double)r.ax, (double)r.ay);     // :119
    r.total = r.dist;                                                            // :166 (same points)
  } else {
    // pygame ruleset: uniform float positions (ballenv_pygame.py:468-482, 454-457)
    uint4 hw = dc.reset_head(episode, 0);
    r.gx = (T)(0.0 + ranf_from_words(hw.x, hw.y) * (cfg.world_w - 0.0));
    r.gy = (T)(0.0 + ranf_from_words(hw.z, hw.w) * (cfg.world_h - 0.0));
    hw = dc.reset_head(episode, 1);
    r.ax = (T)(0.0 + ranf_from_words(hw.x, hw.y) * (cfg.world_w - 0.0));
    r.ay = (T)(0.0 + ranf_from_words(hw.z, hw.w) * (cfg.world_h - 0.0));
    r.dist = dist64((double)r.gx, (double)r.gy, (double)r.ax, (double)r.ay);     // :474, kept even if redrawn (:482)
    for (uint32_t attempt = 0; dist64((double)r.gx, (double)r.gy, (double)r.ax, (double)r.ay) < 50.0; ++attempt) {  // :476-481
      hw = dc.reset_block(episode, (kResetAgentRedraw << 28) | attempt);
      r.ax = (T)(0.0 + ranf_from_words(hw.x, hw.y) * (cfg.world_w - 0.0));
      r.ay = (T)(0.0 + ranf_from_words(hw.z, hw.w) * (cfg.world_h - 0.0));
    }
    r.total = dist64((double)r.ax, (double)r.ay, (double)r.gx, (double)r.gy);    // :511
  }
  return r;
}

// Static obstacle i: redraw until it clears the agent and the goal (ballenv_env.py:131-149 | ballenv_pygame.py:489-498).
template <typename T>
__device__ __forceinline__ void reset_static_draw(const Params& p, const DrawCtx& dc, uint32_t episode, int i,
                                               const ResetHead<T>& h, T& x, T& y) {
  const DevConfig& cfg = p.cfg;
  for (int attempt = 0;; ++attempt) {
    const uint2 w2 = dc.reset_static(episode, i, attempt);
    bool ok;
    if (cfg.ruleset == BALLENV_RULESET_GYM) {
      x = (T)__umulhi(w2.x, 500u);                                               // :24
      y = (T)(20u + __umulhi(w2.y, 460u));                                       // :25
      // check_overlap_rect (:193-197): |dx| < 20 + 5 and |dy| < 20/2 + 5
      const bool ra = fabs((double)x - (double)h.ax) < 25.0 && fabs((double)y - (double)h.ay) < 15.0;
      const bool rg = fabs((double)x - (double)h.gx) < 25.0 && fabs((double)y - (double)h.gy) < 15.0;
      ok = !ra && !rg;
    } else {
      x = (T)__umulhi(w2.x, (uint32_t)cfg.world_w);                              // ballenv_pygame.py:27
      y = (T)__umulhi(w2.y, (uint32_t)cfg.world_h);                              // :32
      const bool oa = !(dist64((double)x, (double)y, (double)h.ax, (double)h.ay) - cfg.reset_agent_thresh > cfg.radius_sum);
      const bool og = !(dist64((double)x, (double)y, (double)h.gx, (double)h.gy) - cfg.reset_goal_thresh > cfg.radius_sum);
      ok = !oa && !og;                                                           // :494
    }
    if (ok) return;
    if (attempt >= kMaxResetAttempts) {
      atomicOr(p.errors, (uint32_t)BALLENV_DEVERR_RESET_STUCK);
      return;
    }
  }
}

// ---- obstacle motion (ballenv_env.py:323-353) --------------------------------------------------------------------
// obstacle move table (:324): (-1,-1) appears twice, (-1,0) is absent.  Packed 2 bits per entry as value + 1.
//   dx: 1 1 1 0 0 0 -1 -1 -1     dy: 1 -1 0 1 -1 0 1 -1 -1
constexpr uint32_t kObstDx = 2u | 2u << 2 | 2u << 4 | 1u << 6 | 1u << 8 | 1u << 10 | 0u << 12 | 0u << 14 | 0u << 16;
constexpr uint32_t kObstDy = 2u | 0u << 2 | 1u << 4 | 2u << 6 | 0u << 8 | 1u << 10 | 2u << 12 | 0u << 14 | 0u << 16;
// agent move table of the training loops (examples/ball_cnn_ac3.py:530)
//   dx: 1 1 1 0 0 0 -1 -1 -1     dy: 1 -1 0 1 -1 0 1 0 -1
constexpr uint32_t kAgentDx = kObstDx;
constexpr uint32_t kAgentDy = 2u | 0u << 2 | 1u << 4 | 2u << 6 | 0u << 8 | 1u << 10 | 2u << 12 | 1u << 14 | 0u << 16;

__device__ __forceinline__ int table2(uint32_t packed, uint32_t i) { return (int)((packed >> (2 * i)) & 3u) - 1; }

template <typename T>
__device__ __forceinline__ void move_obstacle(const DevConfig& cfg, const typename Vec2<T>::type* s_goal, int j,
                                           uint32_t w1, uint32_t w2_tape, bool has_tape, T& x, T& y,
                                           uint32_t& meta) {
  uint32_t gi = meta & 0xffu, cnt = meta >> 8;
  const T s = (T)cfg.speed[j];
  if ((int)cnt < cfg.change_step) {                                   // :327
    const T tx = r_sub(s_goal[gi].x, x), ty = r_sub(s_goal[gi].y, y); // :329-330
    int mx, my;
    if (tx != (T)0 && ty != (T)0) {                                   // :331
      if ((int)__umulhi(w1, 100u) < cfg.rd_th) {                      // :332
        mx = tx > (T)0 ? 1 : -1;                                      // :334-335  tempx / abs(tempx)
        my = ty > (T)0 ? 1 : -1;
      } else {
        const uint32_t w2 = has_tape ? w2_tape : w1 * 100u;           // second draw: unused low half of w1 * 100
        const uint32_t i = __umulhi(w2, 9u);                          // :340
        mx = table2(kObstDx, i);
        my = table2(kObstDy, i);
      }
    } else {
      const uint32_t i = __umulhi(w1, 9u);                            // :345
      mx = table2(kObstDx, i);
      my = table2(kObstDy, i);
    }
    x = r_add(x, r_mul((T)mx, s));
    y = r_add(y, r_mul((T)my, s));
    cnt += 1;                                                         // :348
  } else {                                                            // :349-353 pick another goal, do not move
    if (cfg.goals_distinct) {
      const uint32_t m = __umulhi(w1, (uint32_t)(cfg.n_goals - 1));
      gi = m + (m >= gi ? 1u : 0u);
    } else {
      const T cx = s_goal[gi].x, cy = s_goal[gi].y;
      int others = 0;
      for (int k = 0; k < cfg.n_goals; ++k) others += (s_goal[k].x != cx || s_goal[k].y != cy) ? 1 : 0;
      int m = (int)__umulhi(w1, (uint32_t)others);
      for (int k = 0; k < cfg.n_goals; ++k) {
        if (s_goal[k].x != cx || s_goal[k].y != cy) {
          if (m == 0) {
            gi = (uint32_t)k;
            break;
          }
          --m;
        }
      }
    }
    cnt = 0;
  }
  meta = gi | (cnt << 8);
}

// Hot-path form of move_obstacle for distinct goals: select instead of branch; x + m * s as one FMA, which
// rounds like r_add(x, r_mul(m, s)) because m is -1, 0 or 1 and the product is therefore exact.
template <typename T, int W, bool kSelects = false>
__device__ __forceinline__ void move_lean(const DevConfig& cfg, const BlockShared<T, W>& sh, int j, uint32_t w1,
                                          uint32_t w2, T& x, T& y, uint32_t& meta) {
  const uint32_t gi = meta & 0xffu;
  const typename Vec2<T>::type gl = sh.goal[gi];
  const T tx = r_sub(gl.x, x), ty = r_sub(gl.y, y);                    // :329-330
  const bool diag = tx != (T)0 && ty != (T)0;                          // :331
  const bool seek = diag && (int)__umulhi(w1, 100u) < cfg.rd_th;       // :332
  const uint32_t i = __umulhi(diag ? w2 : w1, 9u);                     // :340 / :345
  const typename Vec2<T>::type mv = sh.mv[i];
  T mx = mv.x, my = mv.y;
  if (seek) {                                                          // :334-335  tempx / abs(tempx)
    mx = tx > (T)0 ? (T)1 : (T)-1;
    my = ty > (T)0 ? (T)1 : (T)-1;
  }
  if constexpr (kSelects) {
    // selects only: measured faster when the quad is partly filled (the per-slot bound tests already branch)
    const bool moving = (int)(meta >> 8) < cfg.change_step;            // :327
    const T s = moving ? CfgV<T>::speed(cfg, j) : (T)0;                // x + m * 0 == x exactly (m is -1, 0 or 1)
    x = fma(mx, s, x);
    y = fma(my, s, y);
    const uint32_t m = __umulhi(w1, (uint32_t)(cfg.n_goals - 1));      // :349-353 pick another goal, do not move
    meta = moving ? meta + 256u : m + (m >= gi ? 1u : 0u);             // :348 / :352-353
  } else if ((int)(meta >> 8) < cfg.change_step) {                     // :327
    const T s = CfgV<T>::speed(cfg, j);
    x = fma(mx, s, x);
    y = fma(my, s, y);
    meta += 256u;                                                      // :348
  } else {                                                             // :349-353 pick another goal, do not move
    const uint32_t m = __umulhi(w1, (uint32_t)(cfg.n_goals - 1));
    meta = m + (m >= gi ? 1u : 0u);
  }
}

// ---- observation store ------------------------------------------------------------------------------------------
// The block's output span [e0, e0 + cnt) x (4 + W*W) elements is contiguous in global memory, starts 128-byte
// aligned (e0 is a multiple of 32), and element f of it is bit f of the block's bit-stream.  Four consecutive
// elements are therefore one aligned nibble: a 16-entry float4 table turns it into one 128-bit streaming store.
__device__ __forceinline__ uint32_t stream_bits(const uint32_t* words, int bit, int nbits_total_words) {
  // 32 bits of the stream starting at `bit` (reads at most one word past the start word, clamped)
  const int wi = bit >> 5, sh = bit & 31;
  const uint32_t lo = words[wi], hi = (wi + 1 < nbits_total_words) ? words[wi + 1] : 0u;
  return __funnelshift_r(lo, hi, sh);
}

// kThreads storing threads, a multiple of 8: thread tid always expands nibble (tid & 7) of the words tid / 8 +
// k * kThreads / 8, so the shift is loop-invariant: rotate the nibble to bits 4..7 (= its byte offset in the
// table), mask, load the float4, store.  A full block (32 environments) of a compile-time window unrolls completely.
template <bool kStream>
__device__ __forceinline__ void put_row4(float4* d, const float4& v) {
  if (kStream) __stcs(d, v);   // global rollout buffer: not re-read by this kernel
  else *d = v;                 // shared-memory staging
}

template <int W, int kThreads, bool kStream = true>
__device__ __forceinline__ void store_rows_f32(float4* dst, const uint32_t* words, const float4* lut, int nvec,
                                               int tid) {
  static_assert(kThreads % 8 == 0, "the nibble a thread expands must not depend on the iteration");
  const uint32_t rot = (((uint32_t)tid & 7u) * 4u + 28u) & 31u;
  const uint32_t* wp = words + (tid >> 3);
  const char* lutb = reinterpret_cast<const char*>(lut);
  float4* d = dst + tid;
  constexpr int kFullVec = W > 0 ? 8 * (4 + W * W) : 0;
  if (W > 0 && nvec == kFullVec) {
    constexpr int kIter = kFullVec / kThreads, kTail = kFullVec % kThreads;
    // batches of four: the word loads, then the table loads, then the stores - four independent chains in flight
    // instead of one (the storing threads are on the step's critical path)
#pragma unroll
    for (int k0 = 0; k0 < kIter; k0 += 4) {
      uint32_t wd[4];
      float4 v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (k0 + j < kIter) wd[j] = wp[(k0 + j) * (kThreads / 8)];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (k0 + j < kIter) v[j] = *reinterpret_cast<const float4*>(lutb + (__funnelshift_r(wd[j], wd[j], rot) & 0xf0u));
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (k0 + j < kIter) put_row4<kStream>(d + (k0 + j) * kThreads, v[j]);
    }
    if (kTail != 0 && tid < kTail) {
      const uint32_t wd = wp[kIter * (kThreads / 8)];
      put_row4<kStream>(d + kIter * kThreads,
                        *reinterpret_cast<const float4*>(lutb + (__funnelshift_r(wd, wd, rot) & 0xf0u)));
    }
  } else {
    for (int v = tid; v < nvec; v += kThreads, wp += kThreads / 8, d += kThreads) {
      const uint32_t wd = *wp;
      put_row4<kStream>(d, *reinterpret_cast<const float4*>(lutb + (__funnelshift_r(wd, wd, rot) & 0xf0u)));
    }
  }
}

// Measured (tools/ab3.sh): the specialised store gains 4 % for WINDOW = 10 rows (13 stores per thread with 64
// threads) and loses 3 % for WINDOW = 5 (2 per thread) - used from three stores per thread on.
template <int W, int kThreads>
__device__ __forceinline__ constexpr bool row_store_pays() {
  return W == 0 || 8 * (4 + W * W) >= 3 * kThreads;
}

// blk points at the block's span of the step's output (environment e0, element 0).
// tid / nthreads: the caller's index among the threads that share the store (the obstacle threads in the step
// loop - the scalar warp is the critical path of a step and stays out of it; everybody after a reset).
// kHot: the call of the step loop (the specialised row store pays there); the stores after a reset stay generic.
template <int W, bool kFast, bool kHot = false>
__device__ __forceinline__ void store_obs(const Params& p, void* blk, const uint32_t* words, const float4* lut, int cnt,
                                          int tid, int nthreads) {
  const int w = Win<W>::w(p.cfg.window);
  const int nb = 4 + w * w;
  const int total = cnt * nb;
  if (kFast || p.cfg.obs_format == BALLENV_OBS_F32) {
    // 16-byte alignment holds unless a [T][n][row] rollout buffer has n * row not a multiple of 4 elements
    const int nvec = (reinterpret_cast<uintptr_t>(blk) & 15) == 0 ? total >> 2 : 0;
    float4* dst = reinterpret_cast<float4*>(blk);
    if (kHot && nthreads == 64 && row_store_pays<W, 64>()) store_rows_f32<W, 64>(dst, words, lut, nvec, tid);
    else if (kHot && nthreads == 128 && row_store_pays<W, 128>()) store_rows_f32<W, 128>(dst, words, lut, nvec, tid);
    else if (kHot && nthreads == kLaneThreads && row_store_pays<W, kLaneThreads>())
      store_rows_f32<W, kLaneThreads>(dst, words, lut, nvec, tid);
    else
    for (int v = tid; v < nvec; v += nthreads)   // streaming store: the rollout buffer is not re-read by this kernel
      __stcs(dst + v, lut[(words[v >> 3] >> ((v & 7) << 2)) & 15u]);
    for (int f = (nvec << 2) + tid; f < total; f += nthreads)   // odd tail of the last block / unaligned span
      reinterpret_cast<float*>(blk)[f] = (words[f >> 5] >> (f & 31)) & 1u ? 1.0f : 0.0f;
  } else if (p.cfg.obs_format == BALLENV_OBS_U8) {
    const int nvec = (reinterpret_cast<uintptr_t>(blk) & 3) == 0 ? total >> 2 : 0;
    uint32_t* dst = reinterpret_cast<uint32_t*>(blk);
    for (int v = tid; v < nvec; v += nthreads) {
      const uint32_t nib = (words[v >> 3] >> ((v & 7) << 2)) & 15u;
      dst[v] = (nib & 1u) | (nib & 2u) << 7 | (nib & 4u) << 14 | (nib & 8u) << 21;
    }
    for (int f = (nvec << 2) + tid; f < total; f += nthreads)
      reinterpret_cast<uint8_t*>(blk)[f] = (uint8_t)((words[f >> 5] >> (f & 31)) & 1u);
  } else {  // BALLENV_OBS_BITS: [n][ceil(nb / 32)] words, bit b of row e = element b of environment e
    const int nw = (nb + 31) >> 5;
    uint32_t* base = reinterpret_cast<uint32_t*>(blk);
    for (int i = tid; i < cnt * nw; i += nthreads) {
      const int en = i / nw, k = i - en * nw;
      uint32_t v = stream_bits(words, en * nb + 32 * k, nb);
      const int left = nb - 32 * k;
      if (left < 32) v &= (1u << left) - 1u;
      base[i] = v;
    }
  }
}

__device__ __forceinline__ int goal_quadrant_bit(bool dx_neg, bool dy_neg) {
  // prep_state2 (examples/ball_cnn_ac3.py:341-350): idx1 dx>=0,dy>=0 ; idx0 dx<0,dy>=0 ; idx3 dx<0,dy<0 ; idx2 else
  // as a packed table over (dx_neg | dy_neg << 1): 1, 0, 2, 3
  return (int)((0xe1u >> (2u * ((dx_neg ? 1u : 0u) | (dy_neg ? 2u : 0u)))) & 3u);
}

// Dynamic quad at element offset off: 128-bit loads of x, y (and meta when stepping) ...
template <typename T>
__device__ __forceinline__ void dynamic_load(const Params& p, uint32_t off, bool stepping, T (&x)[4], T (&y)[4],
                                             uint32_t (&meta)[4]) {
  load4(reinterpret_cast<const T*>(p.dyn_x) + off, x);
  load4(reinterpret_cast<const T*>(p.dyn_y) + off, y);
  if (stepping) load4(p.dyn_meta + off, meta);
}

// ... then one Philox block and four moves, in registers (ballenv_env.py:262-264, 323-353) ...
// kFast: Philox draws (no tape) and distinct goals - the production case, with the uniform tests hoisted out of
// the per-obstacle code.
// kFull (with kFast): all four slots of the quad hold obstacles, no per-slot bound test.
template <typename T, int W, bool kFast, bool kFull = false>
__device__ __forceinline__ void dynamic_move(const Params& p, const BlockShared<T, W>& sh, long long e, int jq,
                                             uint32_t tick, T (&x)[4], T (&y)[4], uint32_t (&meta)[4]) {
  const DevConfig& cfg = p.cfg;
  const bool has_tape = !kFast && p.step_tape != nullptr;
  uint4 blk = make_uint4(0, 0, 0, 0);
  if (!has_tape) blk = philox4x32_10(p.g0 + (uint32_t)e, tick, (uint32_t)jq, kStreamStep, p.k0, p.k1);
  if (kFast && kFull) {
    // The step counters of an environment's obstacles run in lockstep (all start at 0 with the episode, all pick a
    // new goal on the same step), so the quad almost always moves as a whole: one branch for it instead of four,
    // and four independent chains inside.  meta = goal | count << 8, so count < change_step <=> meta < change_step << 8.
    const uint32_t lim = (uint32_t)cfg.change_step << 8;
    const uint32_t top = max(max(meta[0], meta[1]), max(meta[2], meta[3]));
    if (top < lim) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t w1 = pick_word(blk, i);
        const typename Vec2<T>::type gl = sh.goal[meta[i] & 0xffu];
        const T tx = r_sub(gl.x, x[i]), ty = r_sub(gl.y, y[i]);
        const bool diag = tx != (T)0 && ty != (T)0;
        const bool seek = diag && (int)__umulhi(w1, 100u) < cfg.rd_th;
        const typename Vec2<T>::type mv = sh.mv[__umulhi(diag ? w1 * 100u : w1, 9u)];
        const T mx = seek ? (tx > (T)0 ? (T)1 : (T)-1) : mv.x, my = seek ? (ty > (T)0 ? (T)1 : (T)-1) : mv.y;
        const T s = CfgV<T>::speed(cfg, 4 * jq + i);
        x[i] = fma(mx, s, x[i]);
        y[i] = fma(my, s, y[i]);
        meta[i] += 256u;
      }
      return;
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int j = 4 * jq + i;
    if (kFast) {
      const uint32_t w1 = pick_word(blk, i);
      if (kFull) move_lean<T, W, false>(cfg, sh, j, w1, w1 * 100u, x[i], y[i], meta[i]);
      else if (j < cfg.kd) move_lean<T, W, true>(cfg, sh, j, w1, w1 * 100u, x[i], y[i], meta[i]);
    } else if (j < cfg.kd) {
      uint32_t w1 = pick_word(blk, i), w2 = w1 * 100u;   // second draw: unused low half of w1 * 100
      if (has_tape) {
        const uint2 tw = reinterpret_cast<const uint2*>(p.step_tape)[e * cfg.kd + j];
        w1 = tw.x;
        w2 = tw.y;
      }
      if (cfg.goals_distinct) move_lean<T, W>(cfg, sh, j, w1, w2, x[i], y[i], meta[i]);
      else move_obstacle<T>(cfg, sh.goal, j, w1, w2, true, x[i], y[i], meta[i]);
    }
  }
}

// ... and the 128-bit write-back.
template <typename T>
__device__ __forceinline__ void dynamic_store(const Params& p, uint32_t off, const T (&x)[4], const T (&y)[4],
                                              const uint32_t (&meta)[4]) {
  store4(reinterpret_cast<T*>(p.dyn_x) + off, x);
  store4(reinterpret_cast<T*>(p.dyn_y) + off, y);
  store4(p.dyn_meta + off, meta);
}

// slot / per for slot * per < 2^32, per >= 1, with rcp = ceil(2^32 / per) (rcp == 0 encodes per == 1)
__device__ __forceinline__ int div_slot(int slot, uint32_t rcp) {
  return rcp == 0 ? slot : (int)__umulhi((uint32_t)slot, rcp);
}

// ---- stage D of the kernel (rare): (auto-)reset ---------------------------------------------------------------------
// Kept out of line so that its register needs (Philox blocks, rejection loops, fp64 distances) do not count
// against the per-step path.  While it runs the rest of the block waits, so it is organised for a short critical
// path rather than for few instructions:
//   1. the scalar thread of every resetting environment clears the environment's observation bits, draws the head
//      (goal, agent, distances), writes the scalar state and publishes the head through shared memory;
//   2. the obstacle draws of all resetting environments form one item space (environment, obstacle) spread over all
//      256 obstacle threads - one obstacle per thread per pass, each with its own Philox block(s) and rejection loop,
//      so the draw order of the reference (ballenv_env.py:131-164) needs no sequential stream;
//   3. the block rasterises the near list of the new episodes.
// The observation of a finished environment thereby becomes the first observation of its next episode.
template <typename T, int W>
__device__ __noinline__ void reset_stage(const Params& p, BlockShared<T, W>& sh, uint32_t* words, long long e0,
                                         int cnt_env, bool want_obs, int par, int nlt) {
  const DevConfig& cfg = p.cfg;
  const int tid = threadIdx.x;
  const bool is_scalar = tid < 32;
  const int lt = tid - 32;
  const int w = Win<W>::w(cfg.window);
  const int nb = 4 + w * w;
  const int ks = cfg.ks, kd = cfg.kd, K = ks + kd;
  const T margin = CfgV<T>::margin(cfg);

  if (is_scalar) {
    const bool mine = tid < cnt_env && sh.reset[tid] != 0;
    // compact list of the resetting environments
    const uint32_t m = __ballot_sync(0xffffffffu, mine);
    if (mine) sh.rlist[__popc(m & ((1u << tid) - 1u))] = (uint8_t)tid;
    if (tid == 0) {
      sh.nreset = __popc(m);
      sh.count = 0;
    }
    if (mine) {
      for (int b = tid * nb, end = b + nb; b < end;) {   // clear this environment's bits of the stream
        const int sh_ = b & 31, take = min(32 - sh_, end - b);
        const uint32_t mask = (take == 32 ? 0xffffffffu : ((1u << take) - 1u)) << sh_;
        atomicAnd(&words[b >> 5], ~mask);
        b += take;
      }
      const long long e = e0 + tid;
      const DrawCtx dc{&p, e, p.g0 + (uint32_t)e};
      const uint32_t episode = p.episode[e] + 1;
      const ResetHead<T> head = reset_head_draw<T>(p, dc, episode);
      sh.s_gx[tid] = head.gx;
      sh.s_gy[tid] = head.gy;
      sh.s_tick[tid] = episode;
      sh.ax[par][tid] = head.ax;
      sh.ay[par][tid] = head.ay;
      p.episode[e] = episode;
      reinterpret_cast<T*>(p.agent_x)[e] = head.ax;
      reinterpret_cast<T*>(p.agent_y)[e] = head.ay;
      reinterpret_cast<T*>(p.goal_x)[e] = head.gx;
      reinterpret_cast<T*>(p.goal_y)[e] = head.gy;
      p.dist[e] = head.dist;
      p.total[e] = head.total;
      p.acc[e] = 0.0;
      p.ep_len[e] = 0;
      if (want_obs) {
        const T qdx = r_sub(head.gx, head.ax), qdy = r_sub(head.gy, head.ay);
        const int b = tid * nb + goal_quadrant_bit(qdx < (T)0, qdy < (T)0);
        atomicOr(&words[b >> 5], 1u << (b & 31));
      }
    }
  }
  __syncthreads();

  if (!is_scalar && K > 0) {
    T* stat_x = reinterpret_cast<T*>(p.stat_x);
    T* stat_y = reinterpret_cast<T*>(p.stat_y);
    T* dyn_x = reinterpret_cast<T*>(p.dyn_x);
    T* dyn_y = reinterpret_cast<T*>(p.dyn_y);
    // item = (resetting environment r, obstacle k): k fastest, padded to a power of two so that no division is needed
    int kbits = 0;
    while ((1 << kbits) < K) ++kbits;
    const int items = sh.nreset << kbits;
    for (int it = lt; it < items; it += nlt) {
      const int k = it & ((1 << kbits) - 1);
      if (k >= K) continue;
      const int el = sh.rlist[it >> kbits];
      const long long e = e0 + el;
      const DrawCtx dc{&p, e, p.g0 + (uint32_t)e};
      const uint32_t episode = sh.s_tick[el];
      ResetHead<T> head;
      head.gx = sh.s_gx[el];
      head.gy = sh.s_gy[el];
      head.ax = sh.ax[par][el];
      head.ay = sh.ay[par][el];
      head.dist = head.total = 0.0;
      T x, y;
      if (k < ks) {                                                                // :131-149
        reset_static_draw<T>(p, dc, episode, k, head, x, y);
        stat_x[e * p.stat_stride + k] = x;
        stat_y[e * p.stat_stride + k] = y;
      } else {                                                                     // :153-164
        const int j = k - ks;
        const uint2 w2 = dc.reset_dynamic(episode, j);
        x = (T)__umulhi(w2.x, 500u);
        y = (T)(20u + __umulhi(w2.y, 460u));
        dyn_x[e * p.dyn_stride + j] = x;
        dyn_y[e * p.dyn_stride + j] = y;
        p.dyn_meta[e * p.dyn_stride + j] = (uint32_t)j;   // curr_goal = goal_list[j], curr_counter = 0
      }
      near_test<T, W>(sh, words, cfg, el, head.ax, head.ay, margin, x, y, k, false, want_obs, nb);
    }
  }
  __syncthreads();
  if (!is_scalar && want_obs) raster_list<T, W>(sh, words, cfg, lt, nb, par, nlt);
  __syncthreads();
  if (tid == 0) sh.count = 0;   // the list is consumed; the step loop's next pushes come after its kBarAgent
}

// One-step-ahead action fetch for int64 indices: cp.async copies the 8 bytes global -> shared without a register,
// so neither a spill nor an early use can pull the load latency back onto the scalar warp's critical path.
__device__ __forceinline__ void action_fetch_async(long long* smem_slot, const long long* gmem) {
  const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem_slot);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(gmem) : "memory");
}
__device__ __forceinline__ void action_fetch_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ long long load_action_index(const Params& p, long long i) {
  if (p.action_kind == BALLENV_ACT_INDEX_I64) return reinterpret_cast<const long long*>(p.actions)[i];
  if (p.action_kind == BALLENV_ACT_INDEX_I32) return reinterpret_cast<const int*>(p.actions)[i];
  return reinterpret_cast<const uint8_t*>(p.actions)[i];
}

// ---- named barriers (PTX barrier.*, the forms that tolerate intra-warp divergence): the two thread roles below synchronise as producer / consumer ------------------
//   kBarAgent : the scalar warp ARRIVES (does not wait) once the agent positions are published; obstacle
//               threads SYNC on it before their bounding-box tests.  Count = all 288 threads.
//   kBarNear  : everybody syncs: hits and the near list are complete.
//   kBarDone  : the scalar warp publishes "some environment of the block resets" (sh.any_reset) and ARRIVES; it
//               knows the answer from its own vote and runs ahead into the next step (state update, next agent
//               move, outputs) while the obstacle threads, which SYNC here, finish the raster of this one.
//   kBarRaster: split form of kBarDone, used when whole warps hold static quads only (they have no draws and no
//               moves to do).  Every obstacle thread rasterises its share of the near list; the static-quad threads
//               then SYNC here (all rasters done, decision published: count = all 288) and store the rows, while the
//               dynamic-quad threads only ARRIVE here, SYNC on kBarDone (count = 32 + their number) for the decision
//               alone and go straight to their next moves.
//   kBarStore : the storing threads among themselves: the staged rows of the step are complete (bulk store below).
constexpr int kBarAgent = 1, kBarNear = 2, kBarDone = 3, kBarRaster = 4, kBarStore = 5;
template <int kCount = kBlock>
__device__ __forceinline__ void bar_sync(int id) {
  asm volatile("barrier.sync %0, %1;" ::"r"(id), "n"(kCount) : "memory");
}
template <int kCount = kBlock>
__device__ __forceinline__ void bar_arrive(int id) {
  asm volatile("barrier.arrive %0, %1;" ::"r"(id), "n"(kCount) : "memory");
}
__device__ __forceinline__ void bar_sync_n(int id, int count) {
  asm volatile("barrier.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_arrive_n(int id, int count) {
  asm volatile("barrier.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ bool bar_or(int id, bool pred) {
  int r;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %2, 0;\n\tbarrier.red.or.pred q, %1, %3, p;\n\tselp.b32 %0, 1, 0, q;\n\t}"
      : "=r"(r)
      : "r"(id), "r"(pred ? 1 : 0), "n"(kBlock)
      : "memory");
  return r != 0;
}

// ---- bulk store of a block's rows (TMA, cp.async.bulk shared -> global) -----------------------------------------------
// The rows a block produces in one step are one contiguous, 16-byte aligned span of the rollout buffer.  Instead of
// 13 (WINDOW = 10) dependent 128-bit global stores per storing thread, those threads expand the bits into a
// shared-memory staging buffer and one of them hands the whole span to the copy engine; that thread makes sure
// the engine has read the buffer (wait_group.read) before it arrives at the next step's agent barrier, behind which
// the buffer is written again.
__device__ __forceinline__ void bulk_store_rows(void* gmem, const void* smem, uint32_t bytes) {
  const uint32_t src = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem), "r"(src), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_fence_smem_writes() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kPending) : "memory");
}

// ---- the kernel -----------------------------------------------------------------------------------------------------
// A block owns 32 consecutive environments for the whole launch: it loads their state once, advances them
// p.n_steps steps with the state held in registers, and writes it back once.  ballenv_step() launches it with
// n_steps = 1; ballenv_step_many() (open-loop rollouts: the actions of all T steps are given up front) launches it
// once for all T steps, so only actions, observations, rewards and dones touch HBM inside the loop.
//
// Thread roles: warp 0 = scalar warp, thread t owns environment e0 + t (agent move, reward, flags, statistics);
// threads 32 .. 287 = obstacle threads.  The block's obstacle quads form one linear slot space - first the
// 32 * qs static quads, then the 32 * qd dynamic quads, each environment-major exactly as they lie in memory - so
// slot s of a kind sits at element 4 * s of the block's slice, warps are purely static or purely dynamic, and
// consecutive threads issue consecutive 128-bit accesses.  Thread lt keeps slot lt in registers; configurations
// with more than 8 quads per environment go through global memory for the slots beyond 256.
//
// The roles run as two independent instruction streams that meet at named barriers, so neither waits for the
// other's memory latency.  Per step:
//   scalar warp   : action -> move + clamp agent -> publish -> ARRIVE(agent) -> distance, progress reward,
//                   goal / truncation flags (fp64) -> SYNC(near) -> apply hit -> publish reset decision + quadrant
//                   bits -> SYNC(done) -> reward / done out, statistics -> [reset] -> store observations
//   obstacle thr. : (dynamic) Philox, 4 moves -> SYNC(agent) -> clear the other bit-stream buffer -> 4 bounding-box
//                   tests -> SYNC(near) -> raster a share of the near list -> SYNC(done) -> [reset] -> store obs.
//
// kFast: the production specialisation - Step mode, gym ruleset, Philox draws (no tapes), distinct obstacle goals,
// index actions, fp32 observation rows requested.  Every uniform test of the generic kernel folds away; the host
// (ballenv_capi.cu) selects it when all of that holds.
// kRollout: p.n_steps may exceed 1 (ballenv_step_many); otherwise the step loop has exactly one trip and folds away.
// BALLENV_TRACE (tools/phase_trace.py): block 0 writes clock64() stamps of its phases over the reward rows of
// environments 32.. of every step; block 1 keeps its rewards out of the way.  Profiling builds only.
#ifdef BALLENV_TRACE
#define BALLENV_STAMP(cond, slot)                                                                          \
  if (kFast && kRollout && blockIdx.x == 0 && (cond))                                                                         \
    reinterpret_cast<float*>(p.reward)[(long long)t * p.n + 32 + (slot)] = (float)(clock64() & 0xffffff)
#else
#define BALLENV_STAMP(cond, slot)
#endif

// Address of the block's observation rows of step t: (rows of step t) + (rows of the environments before e0).
// Computed where a store needs it, so the step loop carries no pointer.
template <bool kRollout>
__device__ __forceinline__ char* obs_block(const Params& p, long long e0, int t) {
  return reinterpret_cast<char*>(p.obs) +
         ((size_t)(kRollout ? t : 0) * (size_t)p.obs_step_bytes + (size_t)e0 * (size_t)p.obs_row_bytes);
}

__device__ __forceinline__ long long e0_ll(unsigned block) { return (long long)block * kEnvsPerBlock; }

// kSQ = 2 (fast single-step kernel, two or four static quads per environment): a static-quad thread tests TWO quads
// of its environment against the agent and keeps their (constant) coordinates in shared memory instead of
// registers, so the reference's default 13 + 5 obstacles fit blocks of 5 warps (kLW = 4, seven blocks per SM).
// In the rollout loop, and for C3, halving the static-quad threads makes them the long pole of the step (measured
// slower), so the host only selects it for single-step launches of small configurations.
template <typename T, int W, bool kFast, bool kRollout, int kLW = kLanes, int kSQ = 1>
__global__ void __launch_bounds__(32 + 32 * kLW, kMinBlocks<T, kLW>) ballenv_kernel(const __grid_constant__ Params p) {
  constexpr int kLT = 32 * kLW;   // obstacle threads of this instantiation (kLaneThreads = 256 by default)
  constexpr int kB = 32 + kLT;    // block size
  constexpr bool kS2 = kSQ == 2;
  __shared__ __align__(16) T ssx[kS2 ? 64 * 8 : 4], ssy[kS2 ? 64 * 8 : 4];   // static quads of up to 64 threads
  __shared__ BlockShared<T, W> sh;
  // staged rows of a step: rollout kernels with a compile-time window only (13 KB for WINDOW = 10).  One buffer:
  // every kilobyte of shared memory is a kilobyte less L1, and the spilled loop state lives there.
  constexpr int kStageVec = (kFast && kRollout && W > 0 && W <= 10) ? 8 * (4 + W * W) : 0;
  __shared__ __align__(128) float4 stage[kStageVec > 0 ? kStageVec : 1];
  const DevConfig& cfg = p.cfg;
  const int tid = threadIdx.x;
  const long long e0 = (long long)blockIdx.x * kEnvsPerBlock;
  const int cnt_env = (p.n - e0) < kEnvsPerBlock ? (int)(p.n - e0) : kEnvsPerBlock;   // environments of this block
  const int w = Win<W>::w(cfg.window);
  const int nb = 4 + w * w;
  const int ks = cfg.ks, kd = cfg.kd;
  const bool stepping = kFast || p.mode == kModeStep;
  const int n_steps = kRollout ? p.n_steps : 1;
  // threads that store the observation rows: the static-quad threads (no draws, no moves) when they are at least
  // two and not all of the warps, otherwise every obstacle thread; see kBarRaster
  const int n_store = (kRollout && p.n_stat >= (kS2 ? 32 : 64) && p.n_stat < kLT) ? p.n_stat : kLT;
  const bool split = n_store != kLT;
  // bulk store of the rows: full blocks whose span of every step is 16-byte aligned (n_store is 64 or 128 then)
  const bool bulk = kStageVec > 0 && split && (n_store == 32 || n_store == 64 || n_store == 128) && p.n - e0_ll(blockIdx.x) >= kEnvsPerBlock &&
                    ((reinterpret_cast<uintptr_t>(p.obs) + (size_t)e0_ll(blockIdx.x) * (size_t)p.obs_row_bytes) & 15) == 0 &&
                    (p.obs_step_bytes & 15) == 0;
  if (p.debug & 1) return;

  if (tid < 32) {

    // =============================== scalar warp: one thread per environment =====================================
    const bool gym = kFast || cfg.ruleset == BALLENV_RULESET_GYM;
    T* agent_x = reinterpret_cast<T*>(p.agent_x);
    T* agent_y = reinterpret_cast<T*>(p.agent_y);
    T* goal_x = reinterpret_cast<T*>(p.goal_x);
    T* goal_y = reinterpret_cast<T*>(p.goal_y);
    const long long e = e0 + tid;
    const bool mine = tid < cnt_env;
    bool reset_req = false;
    // ---- state of the environment, in registers while the hot loop runs
    T ax = (T)0, ay = (T)0, gx = (T)0, gy = (T)0;
    double dist = 0.0, total = 1.0, acc = 0.0;
    int len = 0;
    uint32_t tick = 0, flags = 0;
    if (mine && !kFast && p.mode == kModeReset) reset_req = p.reset_mask == nullptr || p.reset_mask[e] != 0;
    if (tid < 16)
      sh.lut[tid] = make_float4(tid & 1 ? 1.0f : 0.0f, tid & 2 ? 1.0f : 0.0f, tid & 4 ? 1.0f : 0.0f, tid & 8 ? 1.0f : 0.0f);

    int t = 0, t_fetch = 0;   // t_fetch: the step whose action has to be loaded directly (no fetch in flight)
    bool setup_pending = true;
    for (;;) {
      // (re)load: at launch, and after a reset went through global memory
      if (mine) {
        if (!reset_req) {
          ax = agent_x[e];
          ay = agent_y[e];
          gx = goal_x[e];
          gy = goal_y[e];
        }
        if (stepping) {
          dist = p.dist[e];
          total = p.total[e];
          acc = p.acc[e];
          len = p.ep_len[e];
          tick = p.tick[e];
        }
      }
      if (setup_pending) {   // the state loads above are in flight while the block finishes its setup
        __syncthreads();     // block setup done: the bit-stream of step 0 is cleared before anyone ORs into it
        setup_pending = false;
      }
      // Integral coordinates (what the gym ruleset produces: integer draws, unit steps) stay integral while the
      // loop runs, and the squared distance to the goal is then an exact small integer: sqrt_int22 applies.
      const bool exact32 =
          kFast && __all_sync(0xffffffffu, !mine || (small_int(ax) && small_int(ay) && small_int(gx) && small_int(gy))) &&
          small_int(CfgV<T>::step_x(cfg)) && small_int(CfgV<T>::step_y(cfg)) && small_int(CfgV<T>::world_w(cfg)) &&
          small_int(CfgV<T>::world_h(cfg));
      bool pending_reset = false;
      // Episode statistics (the only thing that is ever all-reduced across GPUs): ballot + shuffle in the scalar
      // warp, one atomic per counter per block, and only in blocks where an episode ended.  cnt: 0, or 1 | 2 goal |
      // 4 static hit | 8 dynamic hit | 16 time-out.
      auto episode_stats = [&](uint32_t cnt, double ret, double ep_len) {
        const uint32_t fin = __ballot_sync(0xffffffffu, cnt != 0);
        double st_ret = cnt ? ret : 0.0, st_len = cnt ? ep_len : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          st_ret += __shfl_xor_sync(0xffffffffu, st_ret, o);
          st_len += __shfl_xor_sync(0xffffffffu, st_len, o);
        }
        const uint32_t n_goal = __popc(__ballot_sync(0xffffffffu, cnt & 2u));
        const uint32_t n_hs = __popc(__ballot_sync(0xffffffffu, cnt & 4u));
        const uint32_t n_hd = __popc(__ballot_sync(0xffffffffu, cnt & 8u));
        const uint32_t n_to = __popc(__ballot_sync(0xffffffffu, cnt & 16u));
        if (tid == 0) {
          atomicAdd(&p.stats[BALLENV_STAT_EPISODES], (double)__popc(fin));
          atomicAdd(&p.stats[BALLENV_STAT_RETURN_SUM], st_ret);
          atomicAdd(&p.stats[BALLENV_STAT_LENGTH_SUM], st_len);
          if (n_goal) atomicAdd(&p.stats[BALLENV_STAT_GOALS], (double)n_goal);
          if (n_hs) atomicAdd(&p.stats[BALLENV_STAT_HITS_STATIC], (double)n_hs);
          if (n_hd) atomicAdd(&p.stats[BALLENV_STAT_HITS_DYNAMIC], (double)n_hd);
          if (n_to) atomicAdd(&p.stats[BALLENV_STAT_TIMEOUTS], (double)n_to);
        }
      };
      // ------------------------------------------- hot loop: no calls inside -------------------------------------
      for (; t < n_steps; ++t) {
        BALLENV_STAMP(tid == 0, 0);
        const bool want_obs = kFast ? (p.obs_all_steps != 0 || t == n_steps - 1) : p.obs != nullptr;
        uint32_t* words = sh.words[t & 1];
        // index of this environment in the [T][n] arrays (n_steps * n < 2^31 per launch: ballenv_step_many splits)
        const long long et = (long long)((uint32_t)t * (uint32_t)p.n + (uint32_t)e);
        // ---- critical: agent move + clamp (ballenv_env.py:247-259 | ballenv_pygame.py:654-664), publish, arrive
        T nx = ax, ny = ay;
        if (mine && stepping) {
          T adx, ady;
          if (!kFast && p.action_kind == BALLENV_ACT_XY_F32) {
            const float2 a = reinterpret_cast<const float2*>(p.actions)[et];
            adx = (T)a.x;
            ady = (T)a.y;
          } else if (!kFast && p.action_kind == BALLENV_ACT_XY_F64) {
            const double2 a = reinterpret_cast<const double2*>(p.actions)[et];
            adx = (T)a.x;
            ady = (T)a.y;
          } else {
            long long ai;
            if (kFast && p.action_kind == BALLENV_ACT_INDEX_I64) {
              // fetched one step ahead into shared memory: the load latency is off the step's critical path
              if (t == t_fetch) {   // first step after a (re)load: nothing in flight yet
                ai = reinterpret_cast<const long long*>(p.actions)[et];
              } else {
                action_fetch_wait();
                ai = sh.act[t & 1][tid];
              }
              if (t + 1 < n_steps)
                action_fetch_async(&sh.act[(t + 1) & 1][tid], reinterpret_cast<const long long*>(p.actions) + et + p.n);
            } else {
              ai = load_action_index(p, et);
            }
            if (ai < 0 || ai > 8) {
              atomicOr(p.errors, (uint32_t)BALLENV_DEVERR_BAD_ACTION);
              ai = 5;  // (0, 0)
            }
            adx = (T)table2(kAgentDx, (uint32_t)ai);
            ady = (T)table2(kAgentDy, (uint32_t)ai);
          }
          if (gym) {
            nx = r_add(ax, r_mul(CfgV<T>::step_x(cfg), adx));   // speedx_ctrl_person * action[0]
            ny = r_add(ay, r_mul(CfgV<T>::step_y(cfg), ady));
          } else {
            nx = r_add(ax, adx);
            ny = r_add(ay, ady);
          }
          if (nx < (T)0) nx = (T)0;
          if (ny < (T)0) ny = (T)0;
          if (nx > CfgV<T>::world_w(cfg)) nx = CfgV<T>::world_w(cfg);
          if (ny > CfgV<T>::world_h(cfg)) ny = CfgV<T>::world_h(cfg);
        }
        sh.ax[t & 1][tid] = nx;
        sh.ay[t & 1][tid] = ny;
        sh.hit[tid] = kNoHit;
        sh.reset[tid] = reset_req ? 1 : 0;   // Reset mode: stored obstacles of these environments are ignored
        if (mine && !reset_req && want_obs) {
          // 4 goal-quadrant bits (examples/ball_cnn_ac3.py:341-350); a reset of this environment clears and redoes them
          const T qdx = r_sub(gx, nx), qdy = r_sub(gy, ny);
          const int b = tid * nb + goal_quadrant_bit(qdx < (T)0, qdy < (T)0);
          atomicOr(&words[b >> 5], 1u << (b & 31));
        }
        BALLENV_STAMP(tid == 0, 1);
        bar_arrive<kB>(kBarAgent);

        // ---- while the obstacle threads move and test: distance, progress reward, goal and time-limit flags of
        //      this step (ballenv_env.py:268-286, 200-206 | ballenv_pygame.py:652-706)
        double d = dist, reward = 0.0;
        bool goal_flag = false, truncated = false;
        const int ep_len = len + 1;
        if (mine && stepping && !(p.debug & 16)) {
          if (exact32) {
            const float dx = (float)gx - (float)nx, dy = (float)gy - (float)ny;     // exact, as are the squares
            d = sqrt_int22(__fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
          } else {
            d = dist64((double)gx, (double)gy, (double)nx, (double)ny);             // :268 | :668
          }
          truncated = cfg.max_steps > 0 && ep_len >= cfg.max_steps;
          goal_flag = d < cfg.goal_threshold;                                      // :276 | :690 (pygame: unless hit)
          if (gym) {
            reward = div64(dist - d, total);                                           // :205-206, old = state[2] (:236)
          } else {
            const double od = dist64((double)ax, (double)ay, (double)gx, (double)gy);   // :652
            reward = div64(od - d, total);                                             // :699-706
          }
        }
        BALLENV_STAMP(tid == 0, 2);
        bar_sync<kB>(kBarNear);
        BALLENV_STAMP(tid == 0, 3);

        // ---- critical: apply the hits, decide the resets
        bool do_reset = false, done_out = false, hit = false, hit_dyn = false, done = false;
        if (mine) {
          if (stepping) {
            const int hit_first = sh.hit[tid];
            hit = hit_first != kNoHit;
            hit_dyn = hit && hit_first >= ks;
            if (gym) {
              if (hit) reward -= hit_dyn ? cfg.dynamic_penalty : cfg.static_penalty;  // :222-224
              done = goal_flag || hit;                                                // :286
            } else {
              if (hit) {                                                              // :683-688 (before the goal test)
                goal_flag = false;
                reward = -1.0;
                done = true;
              } else if (goal_flag) {                                                 // :690-697
                reward = 1.0;
                done = true;
              }
            }
            done_out = done || truncated;
            do_reset = done_out && cfg.auto_reset;
          } else if (!kFast && reset_req) {
            do_reset = true;
          }
          if (do_reset) sh.reset[tid] = 1;
        }
        // The vote is complete inside this warp (it owns all 32 environments): publish it and arrive without
        // waiting.  The scalar warp is the longest dependency chain of a step; everything below, up to the next
        // agent positions, overlaps with the obstacle threads' raster of this step.
        const bool any_reset = __any_sync(0xffffffffu, do_reset);
        if (tid == 0) sh.any_reset = any_reset ? 1 : 0;
        BALLENV_STAMP(tid == 0, 4);
        if (!split) {
          if (any_reset) bar_sync<kB>(kBarDone);   // the reset stage rewrites what the raster of this step still reads
          else bar_arrive<kB>(kBarDone);
        } else {
          bar_arrive_n(kBarDone, 32 + kLT - n_store);
          if (any_reset) bar_sync<kB>(kBarRaster);
          else bar_arrive<kB>(kBarRaster);
        }

        // ---- outputs of the step (the scalar warp is past the barrier: nobody waits for these), next state
        if (mine && stepping) {
          acc += reward;                                                              // :280
          flags = (goal_flag ? BALLENV_FLAG_GOAL : 0) | (hit ? BALLENV_FLAG_HIT : 0) |
                  (truncated ? BALLENV_FLAG_TRUNCATED : 0) | (hit_dyn ? BALLENV_FLAG_HIT_DYNAMIC : 0);
#ifdef BALLENV_TRACE
          if (p.reward != nullptr && blockIdx.x != 1) {
#else
          if (p.reward != nullptr) {
#endif
            if (sizeof(T) == 4) reinterpret_cast<float*>(p.reward)[et] = (float)reward;
            else reinterpret_cast<double*>(p.reward)[et] = reward;
          }
          if (p.done != nullptr) p.done[et] = done_out ? 1 : 0;
          if (e == 0) atomicAdd(&p.stats[BALLENV_STAT_STEPS], (double)p.n);
          ax = nx;
          ay = ny;
          dist = d;
          len = ep_len;
          tick += 1;
        }
        if (stepping && __any_sync(0xffffffffu, done_out)) {   // an episode of the block ended (rare)
          const uint32_t cnt = done_out ? (1u | (goal_flag ? 2u : 0u) | ((hit && !hit_dyn) ? 4u : 0u) |
                                           (hit_dyn ? 8u : 0u) | ((truncated && !done) ? 16u : 0u))
                                        : 0u;
          episode_stats(cnt, acc, (double)ep_len);
        }
        if (any_reset) {   // leave the hot loop: the reset goes through global memory
          pending_reset = true;
          break;
        }
        // (the observation rows of the step are stored by the obstacle threads)
      }

      // ---- write the scalar state back (a pending reset then overwrites it for the environments that finished)
      if (mine) {
        if (stepping) {
          agent_x[e] = ax;
          agent_y[e] = ay;
          p.dist[e] = dist;
          p.acc[e] = acc;
          p.ep_len[e] = len;
          p.tick[e] = tick;
          p.flags[e] = (uint8_t)flags;
        } else if (reset_req) {
          p.flags[e] = 0;
        }
      }
      if (!pending_reset) break;
      {
        // (rare) the step that is being finished is t: reset, then its observation
        const bool want_obs = kFast ? (p.obs_all_steps != 0 || t == n_steps - 1) : p.obs != nullptr;
        uint32_t* words = sh.words[t & 1];
        reset_stage<T, W>(p, sh, words, e0, cnt_env, want_obs, t & 1, kLT);
        if (want_obs && !(p.debug & 8))
          store_obs<W, kFast>(p, obs_block<kRollout>(p, e0, t), words, sh.lut, cnt_env, tid, kB);
        reset_req = false;
        if (++t >= n_steps) break;
      }
    }
  } else {

    // =============================== obstacle threads: one quad per thread per iteration =========================
    const int lt = tid - 32;
    const int qs = (ks + 3) >> 2, qd = (kd + 3) >> 2;         // static / dynamic quads per environment
    const int n_stat = p.n_stat, n_slot = p.n_slot;   // kEnvsPerBlock * qs, kEnvsPerBlock * (qs + qd)
    const T margin = CfgV<T>::margin(cfg);
    // element offsets of the block's obstacle slices (n * K fits 31 bits, checked at create time)
    const uint32_t stat0 = (uint32_t)e0 * (uint32_t)p.stat_stride, dyn0 = (uint32_t)e0 * (uint32_t)p.dyn_stride;
    const T* stat_x = reinterpret_cast<const T*>(p.stat_x);
    const T* stat_y = reinterpret_cast<const T*>(p.stat_y);

    // ---- the thread's own quad (slot lt), in registers while the hot loop runs
    T qx[4], qy[4];
    uint32_t qm[4];
    uint32_t tick = 0;
    int q_el = 0, q_jq = 0;    // environment (of the block), quad of its kind within the environment
    uint32_t q_valid = 0;      // which of the quad's four slots hold obstacles (the last quad of a kind may be padded)
    bool q_have = false;
    const bool q_dyn = lt >= n_stat;
    const uint32_t q_off = q_dyn ? dyn0 + 4u * (uint32_t)(lt - n_stat) : stat0 + (kS2 ? 8u : 4u) * (uint32_t)lt;
    if (lt < n_slot) {
      // slot within its kind; a static-quad thread of a kSQ = 2 kernel owns the quads 2 lt and 2 lt + 1 (qs is even)
      const int sl = q_dyn ? lt - n_stat : (kS2 ? 2 * lt : lt);
      const int per = q_dyn ? qd : qs;
      q_el = div_slot(sl, q_dyn ? cfg.rcp_qd : cfg.rcp_qs);
      const int qq = sl - q_el * per;
      q_have = q_el < cnt_env;
      if (!kFast && q_have && p.mode == kModeReset)
        q_have = p.reset_mask == nullptr ? false : p.reset_mask[e0 + q_el] == 0;
      q_jq = qq;
      const int left = (q_dyn ? kd : ks) - 4 * qq;
      q_valid = left >= 4 ? 15u : (left > 0 ? (1u << left) - 1u : 0u);
      if (kS2 && !q_dyn) q_valid |= (left >= 8 ? 15u : (left > 4 ? (1u << (left - 4)) - 1u : 0u)) << 4;
    }
    // block setup shared by the obstacle threads: cleared bit-stream; goal and move tables.  Every warp writes the
    // (identical) table entries it is going to read, so a warp-level sync is all the moves below need.
    for (int i = lt; i < nb; i += kLT) sh.words[0][i] = 0;
    if (lt == 0) sh.count = 0;
    {
      const int l32 = tid & 31;
      if (l32 < 9) {
        sh.mv[l32].x = (T)table2(kObstDx, (uint32_t)l32);
        sh.mv[l32].y = (T)table2(kObstDy, (uint32_t)l32);
      }
      for (int i = l32; i < cfg.n_goals; i += 32) sh.goal[i] = CfgV<T>::goal(cfg, i);
    }
    int t = 0, t_load = 0;
    bool setup_pending = true;
    for (;;) {
      // (re)load the quad: at launch, and after a reset went through global memory
      t_load = t;
      if (q_have) {
        if (!q_dyn) {
          load4(stat_x + q_off, qx);
          load4(stat_y + q_off, qy);
          if (kS2) {   // both quads go to shared memory: they never change while the loop runs, and nothing stays in registers
            T rx[4], ry[4];
            load4(stat_x + q_off + 4, rx);
            load4(stat_y + q_off + 4, ry);
            store4(ssx + 8 * lt, qx);
            store4(ssy + 8 * lt, qy);
            store4(ssx + 8 * lt + 4, rx);
            store4(ssy + 8 * lt + 4, ry);
          }
        } else {
          if (stepping) tick = p.tick[e0 + q_el] - (uint32_t)t;   // tick of step 0 of this launch
          dynamic_load<T>(p, q_off, stepping, qx, qy, qm);
        }
      }
      if (setup_pending) {   // the quad loads above are in flight while the block finishes its setup
        __syncthreads();     // block setup done (pairs with the scalar warp's)
        setup_pending = false;
      }
      bool pending_reset = false;
      // ------------------------------------------- hot loop: no calls inside -------------------------------------
#ifdef BALLENV_TRACE
      const int tr = lt == 0 ? 8 : (lt == n_stat ? 16 : -1);
#endif
      for (; t < n_steps; ++t) {
        BALLENV_STAMP(tr >= 0, tr + 0);
        const bool want_obs = kFast ? (p.obs_all_steps != 0 || t == n_steps - 1) : p.obs != nullptr;
        uint32_t* words = sh.words[t & 1];
        // obstacle motion does not depend on the agent: draw and move while the scalar warp works
        if (q_have && q_dyn && stepping && !(p.debug & 2)) {
          if (kFast && (kd & 3) == 0)   // block-uniform: every dynamic quad is full
            dynamic_move<T, W, true, true>(p, sh, e0 + q_el, q_jq, tick + (uint32_t)t, qx, qy, qm);
          else if (kFast || (p.step_tape == nullptr && cfg.goals_distinct))
            dynamic_move<T, W, true>(p, sh, e0 + q_el, q_jq, tick + (uint32_t)t, qx, qy, qm);
          else
            dynamic_move<T, W, false>(p, sh, e0 + q_el, q_jq, tick + (uint32_t)t, qx, qy, qm);
        }
        // (the copy engine has had a move phase to read the staged rows of the previous step: free for the next)
        if (kStageVec > 0 && bulk && lt == 0) bulk_wait_read<0>();
        BALLENV_STAMP(tr >= 0, tr + 1);
        bar_sync<kB>(kBarAgent);   // agent positions published; everybody is done with the previous step's bit-stream
        BALLENV_STAMP(tr >= 0, tr + 2);
        if (t + 1 < n_steps)
          for (int i = lt; i < nb; i += kLT) sh.words[(t + 1) & 1][i] = 0;

        // bounding-box test every obstacle against the agent; near ones are hit-tested and queued
        if (kS2 && q_have && !q_dyn && !(p.debug & 4)) {
          const T ax = sh.ax[t & 1][q_el], ay = sh.ay[t & 1][q_el];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            T x[4], y[4];
            load4(ssx + 8 * lt + 4 * h, x);
            load4(ssy + 8 * lt + 4 * h, y);
            uint32_t near = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i)
              near |= (r_abs(r_sub(ax, x[i])) <= margin && r_abs(r_sub(ay, y[i])) <= margin ? 1u : 0u) << i;
            near &= (q_valid >> (4 * h)) & 15u;
            if (near != 0) {
#pragma unroll
              for (int i = 0; i < 4; ++i)
                if (near >> i & 1u)
                  near_push<T, W>(sh, words, cfg, q_el, ax, ay, x[i], y[i], 4 * (q_jq + h) + i, stepping, want_obs, nb);
            }
          }
        } else if (q_have && !(p.debug & 4)) {
          const T ax = sh.ax[t & 1][q_el], ay = sh.ay[t & 1][q_el];
          uint32_t near = 0;   // one branch for the quad: the four tests are almost always all false
#pragma unroll
          for (int i = 0; i < 4; ++i)
            near |= (r_abs(r_sub(ax, qx[i])) <= margin && r_abs(r_sub(ay, qy[i])) <= margin ? 1u : 0u) << i;
          near &= q_valid;
          if (near != 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
              if (near >> i & 1u)
                near_push<T, W>(sh, words, cfg, q_el, ax, ay, qx[i], qy[i], (q_dyn ? ks : 0) + 4 * q_jq + i, stepping,
                                want_obs, nb);
          }
        }
        for (int slot = lt + kLT; slot < n_slot; slot += kLT) {   // more than 8 quads per environment
          const bool dyn = slot >= n_stat;
          const int sl = dyn ? slot - n_stat : slot;
          const int per = dyn ? qd : qs;
          const int el = div_slot(sl, dyn ? cfg.rcp_qd : cfg.rcp_qs), qq = sl - el * per;
          if (el >= cnt_env || sh.reset[el] != 0) continue;
          T x[4], y[4];
          if (!dyn) {
            load4(stat_x + stat0 + 4u * (uint32_t)sl, x);
            load4(stat_y + stat0 + 4u * (uint32_t)sl, y);
          } else {
            uint32_t m[4];
            const uint32_t off = dyn0 + 4u * (uint32_t)sl;
            dynamic_load<T>(p, off, stepping, x, y, m);
            if (stepping) {
              // p.tick is written back only when the hot loop is left: tick of this step = stored + steps since (re)load
              dynamic_move<T, W, false>(p, sh, e0 + el, qq, p.tick[e0 + el] + (uint32_t)(t - t_load), x, y, m);
              dynamic_store<T>(p, off, x, y, m);
            }
          }
          const int k0 = dyn ? ks + 4 * qq : 4 * qq, kend = dyn ? ks + kd : ks;
          const T ax = sh.ax[t & 1][el], ay = sh.ay[t & 1][el];
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (k0 + i < kend)
              near_test<T, W>(sh, words, cfg, el, ax, ay, margin, x[i], y[i], k0 + i, stepping, want_obs, nb);
        }
        BALLENV_STAMP(tr >= 0, tr + 3);
        bar_sync<kB>(kBarNear);
        BALLENV_STAMP(tr >= 0, tr + 4);

        // cooperative raster of the near list, then the reset decision (see kBarRaster)
        if (!split) {
          if (want_obs) raster_list<T, W>(sh, words, cfg, lt, nb, t & 1, kLT);
          BALLENV_STAMP(tr >= 0, tr + 5);
          bar_sync<kB>(kBarDone);
        } else {
          if (want_obs) raster_list<T, W>(sh, words, cfg, lt, nb, t & 1, kLT);   // everybody takes a share of the list
          BALLENV_STAMP(tr >= 0, tr + 5);
          if (lt < n_store) {
            bar_sync<kB>(kBarRaster);    // storing threads: the whole raster and the decision
          } else {
            bar_arrive<kB>(kBarRaster);  // the others: their share is done; they wait for the decision alone
            bar_sync_n(kBarDone, 32 + kLT - n_store);
          }
        }
        BALLENV_STAMP(tr >= 0, tr + 6);
        if (lt == 0) sh.count = 0;   // the near list of the step is consumed (the next pushes come after kBarAgent)
        if (sh.any_reset != 0) {   // leave the hot loop: the reset goes through global memory
          pending_reset = true;
          break;
        }
        // The observation rows are stored by the threads that have the least to do in a step: the static-quad
        // threads (no draws, no moves) when there are at least two warps of them, otherwise every obstacle thread.
        if (want_obs && !(p.debug & 8)) {
          if (kStageVec > 0 && bulk) {
            if (lt < n_store) {
              float4* st = stage;
              if (n_store == 32) store_rows_f32<W, 32, false>(st, words, sh.lut, kStageVec, lt);
              else if (n_store == 64) store_rows_f32<W, 64, false>(st, words, sh.lut, kStageVec, lt);
              else store_rows_f32<W, 128, false>(st, words, sh.lut, kStageVec, lt);
              bulk_fence_smem_writes();
              bar_sync_n(kBarStore, n_store);
              if (lt == 0) {
                bulk_store_rows(obs_block<kRollout>(p, e0, t), st, (uint32_t)(kStageVec * sizeof(float4)));
              }
            }
          } else if (n_store == kLT) {
            store_obs<W, kFast, true>(p, obs_block<kRollout>(p, e0, t), words, sh.lut, cnt_env, lt, kLT);
          } else if (lt < n_store) {
            store_obs<W, kFast, true>(p, obs_block<kRollout>(p, e0, t), words, sh.lut, cnt_env, lt, n_store);
          }
        }
      }
      if (kStageVec > 0 && bulk && lt == 0) bulk_wait_read<0>();   // nothing of this block's staging is still being read

      // ---- write the moved quads back (a pending reset then overwrites those of the environments that finished)
      if (q_have && q_dyn && stepping) dynamic_store<T>(p, q_off, qx, qy, qm);
      if (!pending_reset) break;
      {
        const bool want_obs = kFast ? (p.obs_all_steps != 0 || t == n_steps - 1) : p.obs != nullptr;
        uint32_t* words = sh.words[t & 1];
        reset_stage<T, W>(p, sh, words, e0, cnt_env, want_obs, t & 1, kLT);
        if (want_obs && !(p.debug & 8))
          store_obs<W, kFast>(p, obs_block<kRollout>(p, e0, t), words, sh.lut, cnt_env, tid, kB);
        if (!kFast && lt < n_slot) q_have = q_el < cnt_env;   // Reset mode: the environment has state now
        if (++t >= n_steps) break;
      }
    }
  }
}

}  // namespace ballenv
