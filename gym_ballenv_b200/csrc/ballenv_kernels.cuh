// Fused step + window-observe kernel of the batched ball environment (sm_100a).
//
// One launch advances every environment of a handle by one step:
//   agent move + wall clamp            gym_ballenv/envs/ballenv_env.py:236-259 | ballenv_pygame.py:652-665
//   obstacle motion                    ballenv_env.py:262-264, 323-353
//   distance, goal / obstacle tests,   ballenv_env.py:268-286, 200-229, 179-191 | ballenv_pygame.py:668-706
//   reward, accumulated reward, done
//   TimeLimit(1000) truncation         gym_ballenv/__init__.py:7 (gym 0.10.9 wrapper, restated)
//   auto-reset of finished envs        ballenv_env.py:113-167 | ballenv_pygame.py:460-513 (Philox draws)
//   WINDOW x WINDOW occupancy + goal   examples/ball_cnn_ac3.py:330-352, 384-412 (incl. the row-offset quirk :409)
//   quadrant observation
//
// Mapping: one thread per environment.  State is struct-of-arrays ([K][n] for per-obstacle fields) so every
// state load/store of a warp is one fully used 128-byte line.  The observation rows are produced as bit
// vectors in registers, staged in shared memory and expanded by the whole block into 128-bit coalesced
// stores over the block's contiguous [envs][4 + W*W] output span.  Most obstacles are far from the agent:
// a bounding-box test rejects them in a handful of instructions and only the rare near ones are rasterised
// cell by cell with exactly the reference's arithmetic (dx*dx + dy*dy against the radius sum), so the grid is
// bit-exact by construction.  Nothing here is a dense contraction: no tensor cores.
#pragma once
#include <stdint.h>

#include "../../include/ballenv.h"
#include "ballenv_rng.cuh"

namespace ballenv {

constexpr int kBlock = 128;  // threads (= environments) per block
constexpr int kMaxResetAttempts = 4096;  // the reference would loop forever on an unsatisfiable layout

enum Mode : int { kModeStep = 0, kModeReset = 1, kModeObserve = 2 };

struct DevConfig {
  int ruleset, window, ks, kd, n_goals, goals_distinct, change_step, rd_th;
  int max_steps, auto_reset, obs_format, obs_row_elems;
  double static_penalty, dynamic_penalty;
  double world_w, world_h, radius_sum, goal_threshold, step_x, step_y;
  double reset_agent_thresh, reset_goal_thresh;  // pygame reset clearances (ballenv_pygame.py:494)
  double speed[BALLENV_MAX_DYNAMIC];
  double goal_x[BALLENV_MAX_GOALS], goal_y[BALLENV_MAX_GOALS];
};

struct Params {
  DevConfig cfg;
  long long n, stride;
  uint32_t g0, k0, k1;
  int mode, action_kind;
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

__device__ __forceinline__ double dist64(double ax, double ay, double bx, double by) {
  const double dx = __dsub_rn(ax, bx), dy = __dsub_rn(ay, by);
  return sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
}

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
  static constexpr int kWords = (4 + kMax * kMax + 31) / 32;
  __device__ static __forceinline__ int w(int rt) { return W ? W : rt; }
};

// ---- window raster of one obstacle (examples/ball_cnn_ac3.py:396-409) ----------------------------------------
// Column c samples x = start_x + step_x * c.  Row 0 and row 1 both sample start_y; row r >= 1 samples
// start_y + step_y * (r - 1) because the reference advances cur_y after the column loop with the current r.
template <typename T, int W>
__device__ __forceinline__ void raster_obstacle(T ox, T oy, T start_x, T start_y, T step_x, T step_y, int w,
                                                const Overlap<T>& ov, uint32_t* rows) {
  if constexpr (W == 0) {  // any window size up to 32: plain loops
    for (int r = (w > 1 ? 1 : 0); r < w; ++r) {
      const T dy = r_sub(r_add(start_y, r_mul(step_y, (T)(r > 0 ? r - 1 : 0))), oy);
      uint32_t m = 0;
      for (int c = 0; c < w; ++c) {
        const T dx = r_sub(r_add(start_x, r_mul(step_x, (T)c)), ox);
        m |= (ov(dx, dy) ? 1u : 0u) << c;
      }
      rows[r] |= m;
      if (r <= 1) rows[0] |= m;
    }
  } else {                 // W = 5 / 10: fully unrolled, rows stay in registers
    T dxs[W];
#pragma unroll
    for (int c = 0; c < W; ++c) dxs[c] = r_sub(r_add(start_x, r_mul(step_x, (T)c)), ox);
#pragma unroll
    for (int r = 1; r < W; ++r) {
      const T dy = r_sub(r_add(start_y, r_mul(step_y, (T)(r - 1))), oy);
      uint32_t m = 0;
#pragma unroll
      for (int c = 0; c < W; ++c) m |= (ov(dxs[c], dy) ? 1u : 0u) << c;
      rows[r] |= m;
      if (r == 1) rows[0] |= m;
    }
  }
}

// Per-thread view of one environment while it is being processed.
template <typename T, int W>
struct EnvCtx {
  T ax, ay, gx, gy;          // agent / goal
  T start_x, start_y;        // window origin
  T step_x, step_y;          // cell pitch = agent speed (examples/ball_cnn_ac3.py:390-391)
  T margin;                  // bounding-box half-size beyond which an obstacle cannot touch the window
  uint32_t rows[Win<W>::kMax];
  int hit_first;             // index (list order) of the first obstacle hit, or INT_MAX
};

template <typename T, int W>
__device__ __forceinline__ void ctx_set_agent(EnvCtx<T, W>& c, const DevConfig& cfg, T ax, T ay) {
  const int w = Win<W>::w(cfg.window);
  const int h = w / 2;
  c.ax = ax;
  c.ay = ay;
  c.step_x = (T)cfg.step_x;
  c.step_y = (T)cfg.step_y;
  c.start_x = r_sub(ax, r_mul(c.step_x, (T)h));
  c.start_y = r_sub(ay, r_mul(c.step_y, (T)h));
  const double st = cfg.step_x > cfg.step_y ? cfg.step_x : cfg.step_y;
  c.margin = (T)(cfg.radius_sum + st * (double)h + 2.0);
  if constexpr (W == 0) {
    for (int r = 0; r < w; ++r) c.rows[r] = 0;
  } else {
#pragma unroll
    for (int r = 0; r < W; ++r) c.rows[r] = 0;
  }
  c.hit_first = 0x7fffffff;
}

// Test one obstacle (list index k) against the agent and, if it can touch the window, rasterise it.
template <typename T, int W>
__device__ __forceinline__ void ctx_obstacle(EnvCtx<T, W>& c, const DevConfig& cfg, const Overlap<T>& ov, T ox,
                                             T oy, int k) {
  const T ddx = r_sub(c.ax, ox), ddy = r_sub(c.ay, oy);
  if (r_abs(ddx) <= c.margin && r_abs(ddy) <= c.margin) {
    if (ov(ddx, ddy) && k < c.hit_first) c.hit_first = k;
    raster_obstacle<T, W>(ox, oy, c.start_x, c.start_y, c.step_x, c.step_y, Win<W>::w(cfg.window), ov, c.rows);
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

template <typename T>
__device__ __forceinline__ T* col(void* base, long long k, long long stride) {
  return reinterpret_cast<T*>(base) + k * stride;
}

// ---- reset (ballenv_env.py:113-167 / ballenv_pygame.py:460-513) ------------------------------------------------
// Writes the new obstacle set of env e to global memory, rasterises it into ctx and returns the scalars.
template <typename T, int W>
__device__ __forceinline__ void reset_env(const Params& p, const DrawCtx& dc, const Overlap<T>& ov, uint32_t episode,
                                       EnvCtx<T, W>& c, double& dist, double& total) {
  const DevConfig& cfg = p.cfg;
  const long long e = dc.e, S = p.stride;
  if (cfg.ruleset == BALLENV_RULESET_GYM) {
    const uint4 hw = dc.reset_head(episode, 0);
    const T gx = (T)__umulhi(hw.x, 500u), gy = (T)(480u + __umulhi(hw.y, 20u));  // :115-116
    const T ax = (T)__umulhi(hw.z, 500u), ay = (T)__umulhi(hw.w, 10u);           // :117-118
    // The redraw-while-closer-than-50 loop (:121-126) cannot trigger: goal_y - agent_y >= 471.
    dist = dist64((double)gx, (double)gy, (double)ax, (double)ay);               // :119
    total = dist;                                                                // :166 (same points)
    c.gx = gx;
    c.gy = gy;
    ctx_set_agent<T, W>(c, cfg, ax, ay);
    for (int i = 0; i < cfg.ks; ++i) {                                           // :131-149
      T x, y;
      for (int attempt = 0;; ++attempt) {
        const uint2 w2 = dc.reset_static(episode, i, attempt);
        x = (T)__umulhi(w2.x, 500u);                                             // :24
        y = (T)(20u + __umulhi(w2.y, 460u));                                     // :25
        // check_overlap_rect (:193-197): |dx| < 20 + 5 and |dy| < 20/2 + 5
        const bool ra = fabs((double)x - (double)ax) < 25.0 && fabs((double)y - (double)ay) < 15.0;
        const bool rg = fabs((double)x - (double)gx) < 25.0 && fabs((double)y - (double)gy) < 15.0;
        if (!ra && !rg) break;
        if (attempt >= kMaxResetAttempts) {
          atomicOr(p.errors, (uint32_t)BALLENV_DEVERR_RESET_STUCK);
          break;
        }
      }
      col<T>(p.stat_x, i, S)[e] = x;
      col<T>(p.stat_y, i, S)[e] = y;
      ctx_obstacle<T, W>(c, cfg, ov, x, y, i);
    }
    for (int j = 0; j < cfg.kd; ++j) {                                           // :153-164
      const uint2 w2 = dc.reset_dynamic(episode, j);
      const T x = (T)__umulhi(w2.x, 500u), y = (T)(20u + __umulhi(w2.y, 460u));
      col<T>(p.dyn_x, j, S)[e] = x;
      col<T>(p.dyn_y, j, S)[e] = y;
      p.dyn_meta[(long long)j * S + e] = (uint32_t)j;  // curr_goal = goal_list[j], curr_counter = 0
      ctx_obstacle<T, W>(c, cfg, ov, x, y, cfg.ks + j);
    }
  } else {
    // pygame ruleset: uniform float positions (ballenv_pygame.py:468-482, 454-457)
    uint4 hw = dc.reset_head(episode, 0);
    const double gxd = 0.0 + ranf_from_words(hw.x, hw.y) * (cfg.world_w - 0.0);
    const double gyd = 0.0 + ranf_from_words(hw.z, hw.w) * (cfg.world_h - 0.0);
    hw = dc.reset_head(episode, 1);
    double axd = 0.0 + ranf_from_words(hw.x, hw.y) * (cfg.world_w - 0.0);
    double ayd = 0.0 + ranf_from_words(hw.z, hw.w) * (cfg.world_h - 0.0);
    const T gx = (T)gxd, gy = (T)gyd;
    T ax = (T)axd, ay = (T)ayd;
    dist = dist64((double)gx, (double)gy, (double)ax, (double)ay);               // :474, kept even if redrawn (:482)
    for (uint32_t attempt = 0; dist64((double)gx, (double)gy, (double)ax, (double)ay) < 50.0; ++attempt) {  // :476-481
      hw = dc.reset_block(episode, (kResetAgentRedraw << 28) | attempt);
      ax = (T)(0.0 + ranf_from_words(hw.x, hw.y) * (cfg.world_w - 0.0));
      ay = (T)(0.0 + ranf_from_words(hw.z, hw.w) * (cfg.world_h - 0.0));
    }
    total = dist64((double)ax, (double)ay, (double)gx, (double)gy);              // :511
    c.gx = gx;
    c.gy = gy;
    ctx_set_agent<T, W>(c, cfg, ax, ay);
    for (int i = 0; i < cfg.ks; ++i) {                                           // :489-498
      T x, y;
      for (int attempt = 0;; ++attempt) {
        const uint2 w2 = dc.reset_static(episode, i, attempt);
        x = (T)__umulhi(w2.x, (uint32_t)cfg.world_w);                            // :27
        y = (T)__umulhi(w2.y, (uint32_t)cfg.world_h);                            // :32
        const bool oa = !(dist64((double)x, (double)y, (double)ax, (double)ay) - cfg.reset_agent_thresh > cfg.radius_sum);
        const bool og = !(dist64((double)x, (double)y, (double)gx, (double)gy) - cfg.reset_goal_thresh > cfg.radius_sum);
        if (!oa && !og) break;                                                   // :494
        if (attempt >= kMaxResetAttempts) {
          atomicOr(p.errors, (uint32_t)BALLENV_DEVERR_RESET_STUCK);
          break;
        }
      }
      col<T>(p.stat_x, i, S)[e] = x;
      col<T>(p.stat_y, i, S)[e] = y;
      ctx_obstacle<T, W>(c, cfg, ov, x, y, i);
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
__device__ __forceinline__ void move_obstacle(const DevConfig& cfg, const double* s_goal_x, const double* s_goal_y,
                                              int j, uint32_t w1, uint32_t w2_tape, bool has_tape, T& x, T& y,
                                              uint32_t& meta) {
  uint32_t gi = meta & 0xffu, cnt = meta >> 8;
  const T s = (T)cfg.speed[j];
  if ((int)cnt < cfg.change_step) {                                   // :327
    const T tx = r_sub((T)s_goal_x[gi], x), ty = r_sub((T)s_goal_y[gi], y);   // :329-330
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
      const double cx = s_goal_x[gi], cy = s_goal_y[gi];
      int others = 0;
      for (int k = 0; k < cfg.n_goals; ++k) others += (s_goal_x[k] != cx || s_goal_y[k] != cy) ? 1 : 0;
      int m = (int)__umulhi(w1, (uint32_t)others);
      for (int k = 0; k < cfg.n_goals; ++k) {
        if (s_goal_x[k] != cx || s_goal_y[k] != cy) {
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

// ---- observation store ------------------------------------------------------------------------------------------
// words_s holds, per env of the block, the 4 + W*W observation bits in output order.  The block's output span
// [e0, e0 + cnt) x row is contiguous in global memory and starts 16-byte aligned (e0 is a multiple of 128).
template <int W>
__device__ __forceinline__ void store_obs(const Params& p, const uint32_t* words_s, long long e0, int cnt) {
  constexpr int NWc = Win<W>::kWords;
  const int w = Win<W>::w(p.cfg.window);
  const int nb = 4 + w * w;
  const int nw = W ? NWc : (nb + 31) / 32;
  const int tid = threadIdx.x;
  if (p.cfg.obs_format == BALLENV_OBS_F32) {
    float* base = reinterpret_cast<float*>(p.obs) + e0 * nb;
    const int total = cnt * nb, nvec = total >> 2;
    float4* dst = reinterpret_cast<float4*>(base);
    for (int v = tid; v < nvec; v += kBlock) {
      const int f0 = v * 4;
      float4 o;
      if (W != 0 && ((4 + W * W) % 4 == 0)) {
        const int env = f0 / nb, b0 = f0 - env * nb;  // the 4 bits share a word (b0 % 4 == 0)
        const uint32_t nib = words_s[env * nw + (b0 >> 5)] >> (b0 & 31);
        o.x = (nib & 1u) ? 1.0f : 0.0f;
        o.y = (nib & 2u) ? 1.0f : 0.0f;
        o.z = (nib & 4u) ? 1.0f : 0.0f;
        o.w = (nib & 8u) ? 1.0f : 0.0f;
      } else {
        float t[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int f = f0 + k, env = f / nb, b = f - env * nb;
          t[k] = ((words_s[env * nw + (b >> 5)] >> (b & 31)) & 1u) ? 1.0f : 0.0f;
        }
        o = make_float4(t[0], t[1], t[2], t[3]);
      }
      dst[v] = o;
    }
    for (int f = nvec * 4 + tid; f < total; f += kBlock) {
      const int env = f / nb, b = f - env * nb;
      base[f] = ((words_s[env * nw + (b >> 5)] >> (b & 31)) & 1u) ? 1.0f : 0.0f;
    }
  } else if (p.cfg.obs_format == BALLENV_OBS_U8) {
    uint8_t* base = reinterpret_cast<uint8_t*>(p.obs) + e0 * nb;
    const int total = cnt * nb, nvec = total >> 2;
    uint32_t* dst = reinterpret_cast<uint32_t*>(base);  // e0 * nb is a multiple of 4
    for (int v = tid; v < nvec; v += kBlock) {
      uint32_t o = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int f = v * 4 + k, env = f / nb, b = f - env * nb;
        o |= ((words_s[env * nw + (b >> 5)] >> (b & 31)) & 1u) << (8 * k);
      }
      dst[v] = o;
    }
    for (int f = nvec * 4 + tid; f < total; f += kBlock) {
      const int env = f / nb, b = f - env * nb;
      base[f] = (uint8_t)((words_s[env * nw + (b >> 5)] >> (b & 31)) & 1u);
    }
  } else {  // BALLENV_OBS_BITS
    uint32_t* base = reinterpret_cast<uint32_t*>(p.obs) + e0 * nw;
    for (int i = tid; i < cnt * nw; i += kBlock) base[i] = words_s[i];
  }
}

// ---- the kernel ---------------------------------------------------------------------------------------------------
template <typename T, int W>
__global__ void __launch_bounds__(kBlock) ballenv_kernel(const __grid_constant__ Params p) {
  constexpr int NWc = Win<W>::kWords;
  __shared__ uint32_t words_s[kBlock * NWc];
  __shared__ double s_goal_x[BALLENV_MAX_GOALS], s_goal_y[BALLENV_MAX_GOALS];
  const DevConfig& cfg = p.cfg;
  const int tid = threadIdx.x;
  const long long e0 = (long long)blockIdx.x * kBlock;
  const long long e = e0 + tid;
  const long long S = p.stride;
  const int w = Win<W>::w(cfg.window);
  const int nb = 4 + w * w;
  const int nw = W ? NWc : (nb + 31) / 32;

  if (tid < cfg.n_goals) {
    s_goal_x[tid] = cfg.goal_x[tid];
    s_goal_y[tid] = cfg.goal_y[tid];
  }
  __syncthreads();

  const Overlap<T> ov(cfg.radius_sum);
  // per-thread contribution to the episode statistics (non-zero only for envs that finished this step)
  double st_ret = 0.0, st_len = 0.0;
  uint32_t st_cnt = 0;  // packed one-bit counters: episode | goal << 1 | hit_static << 2 | hit_dynamic << 3 | timeout << 4
  if (e < p.n) {
    T* agent_x = reinterpret_cast<T*>(p.agent_x);
    T* agent_y = reinterpret_cast<T*>(p.agent_y);
    T* goal_x = reinterpret_cast<T*>(p.goal_x);
    T* goal_y = reinterpret_cast<T*>(p.goal_y);
    EnvCtx<T, W> c;
    DrawCtx dc{&p, e, p.g0 + (uint32_t)e};
    bool do_reset = false;
    c.gx = goal_x[e];
    c.gy = goal_y[e];

    if (p.mode == kModeStep) {
      // -------- agent move + clamp (ballenv_env.py:247-259 | ballenv_pygame.py:654-664)
      const T oax = agent_x[e], oay = agent_y[e];
      T adx, ady;
      if (p.action_kind == BALLENV_ACT_XY_F32) {
        const float2 a = reinterpret_cast<const float2*>(p.actions)[e];
        adx = (T)a.x;
        ady = (T)a.y;
      } else if (p.action_kind == BALLENV_ACT_XY_F64) {
        const double2 a = reinterpret_cast<const double2*>(p.actions)[e];
        adx = (T)a.x;
        ady = (T)a.y;
      } else {
        long long ai;
        if (p.action_kind == BALLENV_ACT_INDEX_I64) ai = reinterpret_cast<const long long*>(p.actions)[e];
        else if (p.action_kind == BALLENV_ACT_INDEX_I32) ai = reinterpret_cast<const int*>(p.actions)[e];
        else ai = reinterpret_cast<const uint8_t*>(p.actions)[e];
        if (ai < 0 || ai > 8) {
          atomicOr(p.errors, (uint32_t)BALLENV_DEVERR_BAD_ACTION);
          ai = 5;  // (0, 0)
        }
        adx = (T)table2(kAgentDx, (uint32_t)ai);
        ady = (T)table2(kAgentDy, (uint32_t)ai);
      }
      T nx, ny;
      if (cfg.ruleset == BALLENV_RULESET_GYM) {
        nx = r_add(oax, r_mul((T)cfg.step_x, adx));   // speedx_ctrl_person * action[0]
        ny = r_add(oay, r_mul((T)cfg.step_y, ady));
      } else {
        nx = r_add(oax, adx);
        ny = r_add(oay, ady);
      }
      if (nx < (T)0) nx = (T)0;
      if (ny < (T)0) ny = (T)0;
      if (nx > (T)cfg.world_w) nx = (T)cfg.world_w;
      if (ny > (T)cfg.world_h) ny = (T)cfg.world_h;
      ctx_set_agent<T, W>(c, cfg, nx, ny);

      // -------- obstacles: static (read), dynamic (move, write back); hit test + window raster on the fly
      for (int i = 0; i < cfg.ks; ++i)
        ctx_obstacle<T, W>(c, cfg, ov, col<T>(p.stat_x, i, S)[e], col<T>(p.stat_y, i, S)[e], i);
      const uint32_t tick = p.tick[e];
      const bool has_tape = p.step_tape != nullptr;
      for (int jb = 0; jb < cfg.kd; jb += 4) {
        uint4 blk = make_uint4(0, 0, 0, 0);
        if (!has_tape) blk = philox4x32_10(dc.g, tick, (uint32_t)(jb >> 2), kStreamStep, p.k0, p.k1);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int j = jb + q;
          if (j < cfg.kd) {
            T x = col<T>(p.dyn_x, j, S)[e], y = col<T>(p.dyn_y, j, S)[e];
            uint32_t meta = p.dyn_meta[(long long)j * S + e];
            uint32_t w1 = pick_word(blk, q), w2 = 0;
            if (has_tape) {
              const uint2 tw = reinterpret_cast<const uint2*>(p.step_tape)[e * cfg.kd + j];
              w1 = tw.x;
              w2 = tw.y;
            }
            move_obstacle<T>(cfg, s_goal_x, s_goal_y, j, w1, w2, has_tape, x, y, meta);
            col<T>(p.dyn_x, j, S)[e] = x;
            col<T>(p.dyn_y, j, S)[e] = y;
            p.dyn_meta[(long long)j * S + e] = meta;
            ctx_obstacle<T, W>(c, cfg, ov, x, y, cfg.ks + j);
          }
        }
      }

      // -------- distance, reward, flags (ballenv_env.py:268-286, 200-229 | ballenv_pygame.py:668-706)
      const double total = p.total[e];
      double acc = p.acc[e];
      const bool hit = c.hit_first != 0x7fffffff;
      const bool hit_dyn = hit && c.hit_first >= cfg.ks;
      double d, reward;
      bool goal_flag, done;
      if (cfg.ruleset == BALLENV_RULESET_GYM) {
        const double old = p.dist[e];                                             // :236
        d = dist64((double)c.gx, (double)c.gy, (double)nx, (double)ny);           // :268
        goal_flag = d < cfg.goal_threshold;                                       // :276
        reward = (old - d) / total;                                               // :205-206
        if (hit) reward -= hit_dyn ? cfg.dynamic_penalty : cfg.static_penalty;    // :222-224
        acc += reward;                                                            // :280
        done = goal_flag || hit;                                                  // :286
      } else {
        const double old = dist64((double)oax, (double)oay, (double)c.gx, (double)c.gy);  // :652
        d = dist64((double)nx, (double)ny, (double)c.gx, (double)c.gy);                   // :668
        goal_flag = false;
        if (hit) {                                                                        // :683-688
          acc += -1.0;
          reward = -1.0;
          done = true;
        } else if (d < cfg.goal_threshold) {                                              // :690-697
          goal_flag = true;
          acc += 1.0;
          reward = 1.0;
          done = true;
        } else {                                                                          // :699-706
          reward = (old - d) / total;
          acc += reward;
          done = false;
        }
      }
      int ep_len = p.ep_len[e] + 1;
      const bool truncated = cfg.max_steps > 0 && ep_len >= cfg.max_steps;
      const bool done_out = done || truncated;
      const uint32_t f = (goal_flag ? BALLENV_FLAG_GOAL : 0) | (hit ? BALLENV_FLAG_HIT : 0) |
                         (truncated ? BALLENV_FLAG_TRUNCATED : 0) | (hit_dyn ? BALLENV_FLAG_HIT_DYNAMIC : 0);
      p.flags[e] = (uint8_t)f;
      p.tick[e] = tick + 1;
      if (p.reward != nullptr) {
        if (sizeof(T) == 4) reinterpret_cast<float*>(p.reward)[e] = (float)reward;
        else reinterpret_cast<double*>(p.reward)[e] = reward;
      }
      if (p.done != nullptr) p.done[e] = done_out ? 1 : 0;

      if (done_out) {
        st_ret = acc;
        st_len = (double)ep_len;
        st_cnt = 1u | (goal_flag ? 2u : 0u) | ((hit && !hit_dyn) ? 4u : 0u) | (hit_dyn ? 8u : 0u) |
                 ((truncated && !done) ? 16u : 0u);
      }
      do_reset = done_out && cfg.auto_reset;
      if (!do_reset) {
        agent_x[e] = nx;
        agent_y[e] = ny;
        p.dist[e] = d;
        p.acc[e] = acc;
        p.ep_len[e] = ep_len;
      }
    } else if (p.mode == kModeReset && (p.reset_mask == nullptr || p.reset_mask[e] != 0)) {
      do_reset = true;
      p.flags[e] = 0;
    } else {
      // observe only: raster the stored state
      ctx_set_agent<T, W>(c, cfg, agent_x[e], agent_y[e]);
      for (int i = 0; i < cfg.ks; ++i)
        ctx_obstacle<T, W>(c, cfg, ov, col<T>(p.stat_x, i, S)[e], col<T>(p.stat_y, i, S)[e], i);
      for (int j = 0; j < cfg.kd; ++j)
        ctx_obstacle<T, W>(c, cfg, ov, col<T>(p.dyn_x, j, S)[e], col<T>(p.dyn_y, j, S)[e], cfg.ks + j);
    }

    // -------- (auto-)reset: new episode drawn in place; the observation below is the post-reset one
    if (do_reset) {
      const uint32_t episode = p.episode[e] + 1;
      double rd, rt;
      reset_env<T, W>(p, dc, ov, episode, c, rd, rt);
      p.episode[e] = episode;
      agent_x[e] = c.ax;
      agent_y[e] = c.ay;
      goal_x[e] = c.gx;
      goal_y[e] = c.gy;
      p.dist[e] = rd;
      p.total[e] = rt;
      p.acc[e] = 0.0;
      p.ep_len[e] = 0;
    }

    // -------- pack the observation bits: 4 goal-quadrant bits (examples/ball_cnn_ac3.py:341-350) + W*W cells
    if (p.obs != nullptr) {
      const T qdx = r_sub(c.gx, c.ax), qdy = r_sub(c.gy, c.ay);
      const int q = (qdx >= (T)0 && qdy >= (T)0) ? 1 : ((qdx < (T)0 && qdy >= (T)0) ? 0 : ((qdx < (T)0 && qdy < (T)0) ? 3 : 2));
      if constexpr (W == 0) {
        for (int i = 0; i < nw; ++i) words_s[tid * nw + i] = 0;
        words_s[tid * nw] = 1u << q;
        for (int r = 0; r < w; ++r) {
          const int off = 4 + r * w;
          const uint32_t v = c.rows[r];
          words_s[tid * nw + (off >> 5)] |= v << (off & 31);
          if ((off & 31) + w > 32) words_s[tid * nw + (off >> 5) + 1] |= v >> (32 - (off & 31));
        }
      } else {
        uint32_t words[NWc];
#pragma unroll
        for (int i = 0; i < NWc; ++i) words[i] = 0;
        words[0] = 1u << q;
#pragma unroll
        for (int r = 0; r < W; ++r) {
          const int off = 4 + r * W;
          const uint32_t v = c.rows[r];
          words[off >> 5] |= v << (off & 31);
          if ((off & 31) + W > 32) words[(off >> 5) + 1] |= v >> (32 - (off & 31));
        }
#pragma unroll
        for (int i = 0; i < NWc; ++i) words_s[tid * NWc + i] = words[i];
      }
    }
  }
  if (p.obs != nullptr) {
    __syncthreads();
    const long long rem = p.n - e0;
    store_obs<W>(p, words_s, e0, rem < kBlock ? (int)rem : kBlock);
  }
  if (p.mode == kModeStep) {
    // episode statistics (the only thing that is ever all-reduced across GPUs): ballot + shuffle per warp,
    // one atomic per counter per warp, and only in warps where an episode ended.
    const uint32_t fin = __ballot_sync(0xffffffffu, st_cnt != 0);
    if (fin != 0) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        st_ret += __shfl_xor_sync(0xffffffffu, st_ret, o);
        st_len += __shfl_xor_sync(0xffffffffu, st_len, o);
      }
      const uint32_t n_goal = __popc(__ballot_sync(0xffffffffu, st_cnt & 2u));
      const uint32_t n_hs = __popc(__ballot_sync(0xffffffffu, st_cnt & 4u));
      const uint32_t n_hd = __popc(__ballot_sync(0xffffffffu, st_cnt & 8u));
      const uint32_t n_to = __popc(__ballot_sync(0xffffffffu, st_cnt & 16u));
      if ((tid & 31) == 0) {
        atomicAdd(&p.stats[BALLENV_STAT_EPISODES], (double)__popc(fin));
        atomicAdd(&p.stats[BALLENV_STAT_RETURN_SUM], st_ret);
        atomicAdd(&p.stats[BALLENV_STAT_LENGTH_SUM], st_len);
        if (n_goal) atomicAdd(&p.stats[BALLENV_STAT_GOALS], (double)n_goal);
        if (n_hs) atomicAdd(&p.stats[BALLENV_STAT_HITS_STATIC], (double)n_hs);
        if (n_hd) atomicAdd(&p.stats[BALLENV_STAT_HITS_DYNAMIC], (double)n_hd);
        if (n_to) atomicAdd(&p.stats[BALLENV_STAT_TIMEOUTS], (double)n_to);
      }
    }
    if (e == 0) atomicAdd(&p.stats[BALLENV_STAT_STEPS], (double)p.n);
  }
}

}  // namespace ballenv
