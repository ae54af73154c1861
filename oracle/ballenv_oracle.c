/*
 * ballenv_oracle.c - C restatement of the CPU oracle (gym ruleset, Philox-addressed draws).
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may build, load or call
 * this; the product (gym_ballenv_b200/) never does and has no CPU fallback.
 *
 * Same algorithm, same operation order and same fp64 arithmetic as oracle/ballenv_oracle.py (which is pinned to
 * the reference by tests/golden/, see tests/test_oracle_vs_golden.py); tests/test_oracle_c.py checks this file
 * against the Python restatement bit for bit.  It exists so that parity can be checked at BASELINE.json's full
 * sizes (4096 envs x 200 steps, 65536 envs) in seconds.
 *
 * Reference citations (relative to the reference root):
 *   step            gym_ballenv/envs/ballenv_env.py:232-289
 *   reward / hit    gym_ballenv/envs/ballenv_env.py:200-229, 179-191
 *   obstacle motion gym_ballenv/envs/ballenv_env.py:323-353
 *   reset           gym_ballenv/envs/ballenv_env.py:113-167, 19-33, 193-197
 *   window obs      examples/ball_cnn_ac3.py:330-352 (goal quadrant), 384-412 (W x W raster, quirk at :409)
 *   time limit      gym_ballenv/__init__.py:7 (gym 0.10.9 TimeLimit, third party: parity unpinned, restated)
 *   draws           oracle/draws.py (Philox4x32-10, Salmon et al. SC'11; randint(n) = mulhi(word, n))
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#define ORC_MAX_DYNAMIC 64
#define ORC_MAX_GOALS 64

typedef struct OrcConfig {
  int32_t window, n_static, n_dynamic, n_goals, change_step, rd_th_obs, max_episode_steps, auto_reset;
  double static_penalty, dynamic_penalty;
  double speeds[ORC_MAX_DYNAMIC];
  double goal_x[ORC_MAX_GOALS], goal_y[ORC_MAX_GOALS];
} OrcConfig;

typedef struct OrcEnv {
  double ax, ay, gx, gy, dist, total, acc;
  int32_t ep_len, episode;
  uint32_t tick;
  double *ox, *oy;      /* [K] static first, then dynamic */
  int32_t *goal_idx, *counter; /* [Kd] */
  int goal_flag, hit, hit_index, truncated;
} OrcEnv;

typedef struct OrcVec {
  OrcConfig cfg;
  int64_t n, g0;
  uint32_t k0, k1;
  OrcEnv* env;
  double stats[8]; /* episodes return_sum length_sum goals hits_static hits_dynamic timeouts steps */
} OrcVec;

/* ---- draws (oracle/draws.py) ------------------------------------------------------------------------------- */
enum { STREAM_STEP = 1, STREAM_RESET = 2 };
enum { RK_HEAD = 0, RK_STATIC = 1, RK_DYNAMIC = 2, RK_AGENT_REDRAW = 3 };

static void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static uint32_t mulhi(uint32_t w, uint32_t n) { return (uint32_t)(((uint64_t)w * n) >> 32); }

static void reset_words(const OrcVec* v, uint32_t g, uint32_t episode, int kind, int item, int attempt, uint32_t w[4]) {
  uint32_t c2 = 0, b[4];
  int w0 = 0;
  switch (kind) {
    case RK_HEAD: c2 = ((uint32_t)RK_HEAD << 28) | (uint32_t)item; break;
    case RK_STATIC: c2 = ((uint32_t)RK_STATIC << 28) | ((uint32_t)item << 16) | ((uint32_t)attempt >> 1); w0 = (attempt & 1) * 2; break;
    case RK_DYNAMIC: c2 = ((uint32_t)RK_DYNAMIC << 28) | ((uint32_t)item >> 1); w0 = (item & 1) * 2; break;
    default: c2 = ((uint32_t)RK_AGENT_REDRAW << 28) | (uint32_t)attempt; break;
  }
  philox4x32_10(g, episode, c2, STREAM_RESET, v->k0, v->k1, b);
  for (int i = 0; i + w0 < 4; ++i) w[i] = b[w0 + i];
}

/* ---- geometry ---------------------------------------------------------------------------------------------- */
static double dist2(double px, double py, double qx, double qy) {
  /* calculate_distance, ballenv_env.py:179-183 */
  return sqrt(pow(px - qx, 2) + pow(py - qy, 2));
}

static int rect_overlap(double x, double y, double px, double py) {
  /* check_overlap_rect, ballenv_env.py:193-197 with rad = 20, agent radius 5 */
  return fabs(x - px) < (20 + 5) && fabs(y - py) < (20 / 2.0 + 5);
}

/* ---- reset (ballenv_env.py:113-167) ------------------------------------------------------------------------- */
static void env_reset(const OrcVec* v, OrcEnv* e, uint32_t g) {
  const OrcConfig* c = &v->cfg;
  uint32_t w[4];
  e->episode += 1;
  e->ep_len = 0;
  const uint32_t ep = (uint32_t)e->episode;
  reset_words(v, g, ep, RK_HEAD, 0, 0, w);
  e->gx = mulhi(w[0], 500); e->gy = 480 + mulhi(w[1], 20);            /* :115-116 */
  e->ax = mulhi(w[2], 500); e->ay = mulhi(w[3], 10);                  /* :117-118 */
  e->dist = dist2(e->gx, e->gy, e->ax, e->ay);                       /* :119 */
  for (int attempt = 0; dist2(e->gx, e->gy, e->ax, e->ay) < 50; ++attempt) {   /* :121-126 (dead for this geometry) */
    reset_words(v, g, ep, RK_AGENT_REDRAW, 0, attempt, w);
    e->ax = mulhi(w[0], 500); e->ay = mulhi(w[1], 10);
  }
  e->acc = 0.0;
  for (int i = 0; i < c->n_static; ++i) {                             /* :131-149 */
    for (int attempt = 0;; ++attempt) {
      reset_words(v, g, ep, RK_STATIC, i, attempt, w);
      const double x = mulhi(w[0], 500), y = 20 + mulhi(w[1], 460);   /* :24-25 */
      if (!rect_overlap(x, y, e->ax, e->ay) && !rect_overlap(x, y, e->gx, e->gy)) {
        e->ox[i] = x; e->oy[i] = y;
        break;
      }
    }
  }
  for (int j = 0; j < c->n_dynamic; ++j) {                            /* :153-164 */
    reset_words(v, g, ep, RK_DYNAMIC, j, 0, w);
    e->ox[c->n_static + j] = mulhi(w[0], 500);
    e->oy[c->n_static + j] = 20 + mulhi(w[1], 460);
    e->goal_idx[j] = j;
    e->counter[j] = 0;
  }
  e->total = dist2(e->ax, e->ay, e->gx, e->gy);                      /* :166 */
}

/* ---- obstacle motion (ballenv_env.py:323-353) ---------------------------------------------------------------- */
static const int kMoveX[9] = {1, 1, 1, 0, 0, 0, -1, -1, -1};
static const int kMoveY[9] = {1, -1, 0, 1, -1, 0, 1, -1, -1};   /* (-1,-1) twice, no (-1,0): :324 */
static const int kAgentX[9] = {1, 1, 1, 0, 0, 0, -1, -1, -1};
static const int kAgentY[9] = {1, -1, 0, 1, -1, 0, 1, 0, -1};   /* examples/ball_cnn_ac3.py:530 */

static void move_obstacle(const OrcVec* v, OrcEnv* e, uint32_t g, int j) {
  const OrcConfig* c = &v->cfg;
  const int k = c->n_static + j;
  const double s = c->speeds[j];
  uint32_t b[4];
  philox4x32_10(g, e->tick, (uint32_t)(j >> 2), STREAM_STEP, v->k0, v->k1, b);
  const uint32_t w1 = b[j & 3];
  if (e->counter[j] < c->change_step) {
    const double tx = c->goal_x[e->goal_idx[j]] - e->ox[k], ty = c->goal_y[e->goal_idx[j]] - e->oy[k];
    if (tx != 0 && ty != 0) {
      if ((int)mulhi(w1, 100) < c->rd_th_obs) {
        e->ox[k] += (tx / fabs(tx)) * s;
        e->oy[k] += (ty / fabs(ty)) * s;
      } else {
        const uint32_t w2 = (uint32_t)((uint64_t)w1 * 100u);   /* low half of w1 * 100 */
        const int m = (int)mulhi(w2, 9);
        e->ox[k] += kMoveX[m] * s;
        e->oy[k] += kMoveY[m] * s;
      }
    } else {
      const int m = (int)mulhi(w1, 9);
      e->ox[k] += kMoveX[m] * s;
      e->oy[k] += kMoveY[m] * s;
    }
    e->counter[j] += 1;
  } else {
    const double cx = c->goal_x[e->goal_idx[j]], cy = c->goal_y[e->goal_idx[j]];
    int others = 0;
    for (int q = 0; q < c->n_goals; ++q) others += (c->goal_x[q] != cx || c->goal_y[q] != cy);
    int m = (int)mulhi(w1, (uint32_t)others);
    for (int q = 0; q < c->n_goals; ++q) {
      if (c->goal_x[q] != cx || c->goal_y[q] != cy) {
        if (m == 0) { e->goal_idx[j] = q; break; }
        --m;
      }
    }
    e->counter[j] = 0;
  }
}

/* ---- step (ballenv_env.py:232-289, 200-229) -> reward, done (no time limit, no auto-reset here) --------------- */
static double env_step(const OrcVec* v, OrcEnv* e, uint32_t g, double adx, double ady, int* done) {
  const OrcConfig* c = &v->cfg;
  const int K = c->n_static + c->n_dynamic;
  const double old = e->dist;                                   /* :236 */
  double nx = e->ax + 1 * adx, ny = e->ay + 1 * ady;            /* :247-250 */
  if (nx < 0) nx = 0;
  if (ny < 0) ny = 0;
  if (nx > 500) nx = 500;
  if (ny > 500) ny = 500;
  for (int j = 0; j < c->n_dynamic; ++j) move_obstacle(v, e, g, j);   /* :262-264 */
  e->ax = nx; e->ay = ny;
  e->dist = dist2(e->gx, e->gy, e->ax, e->ay);                 /* :268 */
  e->goal_flag = e->dist < 10;                                  /* :276 */
  double reward = -0 + (old - e->dist) / e->total;              /* :205-206 */
  e->hit = 0; e->hit_index = -1;
  for (int k = 0; k < K; ++k) {                                 /* :208-224 */
    if (!(dist2(e->ax, e->ay, e->ox[k], e->oy[k]) > 25)) {
      e->hit = 1; e->hit_index = k;
      reward -= k < c->n_static ? c->static_penalty : c->dynamic_penalty;
      break;
    }
  }
  e->acc += reward;                                             /* :280 */
  *done = e->goal_flag || e->hit;                               /* :286 */
  e->tick += 1;
  e->ep_len += 1;
  return reward;
}

/* ---- prep_state4 (examples/ball_cnn_ac3.py:384-412) ---------------------------------------------------------- */
static void env_observe(const OrcVec* v, const OrcEnv* e, float* out) {
  const OrcConfig* c = &v->cfg;
  const int W = c->window, K = c->n_static + c->n_dynamic, h = W / 2;
  memset(out, 0, sizeof(float) * (size_t)(4 + W * W));
  const double dx = e->gx - e->ax, dy = e->gy - e->ay;          /* prep_state2, :330-352 */
  const int q = (dx >= 0 && dy >= 0) ? 1 : ((dx < 0 && dy >= 0) ? 0 : ((dx < 0 && dy < 0) ? 3 : 2));
  out[q] = 1.0f;
  const double sx = e->ax - 1 * h, sy = e->ay - 1 * h;
  double y = sy;
  for (int r = 0; r < W; ++r) {
    for (int col = 0; col < W; ++col) {
      const double x = sx + 1 * col;
      for (int k = 0; k < K; ++k) {
        if (!(dist2(x, y, e->ox[k], e->oy[k]) > 25)) { out[4 + r * W + col] = 1.0f; break; }
      }
    }
    y = sy + 1 * r;                                             /* the quirk of :409: advanced with the current r */
  }
}

/* ---- a static-partition parallel-for over environments (pthreads; environments are independent) ------------- */
typedef void (*orc_body)(void* ctx, int64_t i);
typedef struct { orc_body fn; void* ctx; int64_t lo, hi; } OrcJob;
static void* orc_job_main(void* a) {
  OrcJob* j = (OrcJob*)a;
  for (int64_t i = j->lo; i < j->hi; ++i) j->fn(j->ctx, i);
  return NULL;
}
static int orc_threads(void) {
  const char* e = getenv("ORC_THREADS");
  long n = e ? atol(e) : sysconf(_SC_NPROCESSORS_ONLN);
  if (n < 1) n = 1;
  if (n > 64) n = 64;
  return (int)n;
}
static void orc_parallel_for(int64_t n, orc_body fn, void* ctx) {
  int nt = orc_threads();
  if (n < 256 || nt == 1) {
    for (int64_t i = 0; i < n; ++i) fn(ctx, i);
    return;
  }
  pthread_t th[64];
  OrcJob job[64];
  for (int t = 0; t < nt; ++t) {
    job[t].fn = fn; job[t].ctx = ctx; job[t].lo = n * t / nt; job[t].hi = n * (t + 1) / nt;
    pthread_create(&th[t], NULL, orc_job_main, &job[t]);
  }
  for (int t = 0; t < nt; ++t) pthread_join(th[t], NULL);
}

/* ---- vector wrapper (OracleVec of oracle/ballenv_oracle.py) --------------------------------------------------- */
OrcVec* orc_create(const OrcConfig* cfg, int64_t n, uint64_t seed, int64_t g0) {
  OrcVec* v = (OrcVec*)calloc(1, sizeof(OrcVec));
  if (!v) return NULL;
  v->cfg = *cfg; v->n = n; v->g0 = g0;
  v->k0 = (uint32_t)seed; v->k1 = (uint32_t)(seed >> 32);
  v->env = (OrcEnv*)calloc((size_t)n, sizeof(OrcEnv));
  const int K = cfg->n_static + cfg->n_dynamic;
  for (int64_t i = 0; i < n; ++i) {
    OrcEnv* e = &v->env[i];
    e->episode = -1;
    e->total = 1.0;
    e->ox = (double*)calloc((size_t)(K > 0 ? K : 1), sizeof(double));
    e->oy = (double*)calloc((size_t)(K > 0 ? K : 1), sizeof(double));
    e->goal_idx = (int32_t*)calloc((size_t)(cfg->n_dynamic > 0 ? cfg->n_dynamic : 1), sizeof(int32_t));
    e->counter = (int32_t*)calloc((size_t)(cfg->n_dynamic > 0 ? cfg->n_dynamic : 1), sizeof(int32_t));
  }
  return v;
}

void orc_destroy(OrcVec* v) {
  if (!v) return;
  for (int64_t i = 0; i < v->n; ++i) {
    free(v->env[i].ox); free(v->env[i].oy); free(v->env[i].goal_idx); free(v->env[i].counter);
  }
  free(v->env);
  free(v);
}

static void body_reset(void* ctx, int64_t i) {
  OrcVec* v = (OrcVec*)ctx;
  env_reset(v, &v->env[i], (uint32_t)(v->g0 + i));
}
void orc_reset(OrcVec* v) { orc_parallel_for(v->n, body_reset, v); }

typedef struct { OrcVec* v; const int64_t* actions; double* reward; uint8_t* done; uint8_t* flags; float* obs; } OrcStepCtx;
static void body_step(void* ctx, int64_t i) {
  OrcStepCtx* s = (OrcStepCtx*)ctx;
  OrcVec* v = s->v;
  const OrcConfig* c = &v->cfg;
  OrcEnv* e = &v->env[i];
  int d = 0;
  const int a = (int)s->actions[i];
  s->reward[i] = env_step(v, e, (uint32_t)(v->g0 + i), kAgentX[a], kAgentY[a], &d);
  e->truncated = c->max_episode_steps > 0 && e->ep_len >= c->max_episode_steps;
  s->done[i] = (uint8_t)(d || e->truncated);
  s->flags[i] = (uint8_t)((e->goal_flag ? 1 : 0) | (e->hit ? 2 : 0) | (e->truncated ? 4 : 0) |
                          ((e->hit && e->hit_index >= c->n_static) ? 8 : 0));
}
static void body_autoreset(void* ctx, int64_t i) {
  OrcStepCtx* s = (OrcStepCtx*)ctx;
  if (s->done[i]) env_reset(s->v, &s->v->env[i], (uint32_t)(s->v->g0 + i));
}
static void body_observe(void* ctx, int64_t i) {
  OrcStepCtx* s = (OrcStepCtx*)ctx;
  const int row = 4 + s->v->cfg.window * s->v->cfg.window;
  env_observe(s->v, &s->v->env[i], s->obs + (size_t)i * row);
}

/* actions: indices into the agent move table; reward / done / flags per env (flags: 1 goal, 2 hit, 4 truncated, 8 dynamic hit) */
void orc_step(OrcVec* v, const int64_t* actions, double* reward, uint8_t* done, uint8_t* flags) {
  const OrcConfig* c = &v->cfg;
  OrcStepCtx sc = {v, actions, reward, done, flags, NULL};
  orc_parallel_for(v->n, body_step, &sc);
  /* statistics and auto-reset in environment order, like the Python loop (the sums are order-sensitive) */
  for (int64_t i = 0; i < v->n; ++i) {
    OrcEnv* e = &v->env[i];
    v->stats[7] += 1;
    if (done[i]) {
      const int d = e->goal_flag || e->hit;
      v->stats[0] += 1;
      v->stats[1] += e->acc;
      v->stats[2] += e->ep_len;
      v->stats[3] += e->goal_flag ? 1 : 0;
      v->stats[4] += (e->hit && !(flags[i] & 8)) ? 1 : 0;
      v->stats[5] += (flags[i] & 8) ? 1 : 0;
      v->stats[6] += (e->truncated && !d) ? 1 : 0;
    }
  }
  if (c->auto_reset) orc_parallel_for(v->n, body_autoreset, &sc);
}

void orc_observe(OrcVec* v, float* obs) {
  OrcStepCtx sc = {v, NULL, NULL, NULL, NULL, obs};
  orc_parallel_for(v->n, body_observe, &sc);
}

void orc_stats(const OrcVec* v, double out[8]) { memcpy(out, v->stats, sizeof(double) * 8); }

/* state out: scalars [n][7] = ax ay gx gy dist total acc ; ints [n][3] = ep_len episode tick ; obstacles [n][K][2] ;
 * dynamic [n][Kd][2] = goal index, counter.  Any pointer may be NULL. */
void orc_get_state(const OrcVec* v, double* scalars, int64_t* ints, double* obst, int32_t* dyn) {
  const int K = v->cfg.n_static + v->cfg.n_dynamic, Kd = v->cfg.n_dynamic;
  for (int64_t i = 0; i < v->n; ++i) {
    const OrcEnv* e = &v->env[i];
    if (scalars) {
      double* s = scalars + i * 7;
      s[0] = e->ax; s[1] = e->ay; s[2] = e->gx; s[3] = e->gy; s[4] = e->dist; s[5] = e->total; s[6] = e->acc;
    }
    if (ints) { ints[i * 3] = e->ep_len; ints[i * 3 + 1] = e->episode; ints[i * 3 + 2] = e->tick; }
    if (obst) for (int k = 0; k < K; ++k) { obst[(i * K + k) * 2] = e->ox[k]; obst[(i * K + k) * 2 + 1] = e->oy[k]; }
    if (dyn) for (int j = 0; j < Kd; ++j) { dyn[(i * Kd + j) * 2] = e->goal_idx[j]; dyn[(i * Kd + j) * 2 + 1] = e->counter[j]; }
  }
}
