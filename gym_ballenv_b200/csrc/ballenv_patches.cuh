// 40 x 40 rgb patch observation of the pixel policies (sm_100a) - SURVEY 8(f) #4.
//
//   frame (gym viewer, restated)   gym_ballenv/envs/ballenv_env.py:295-314, 357-386 (rendering.make_circle(res=30),
//                                  FilledPolygon, draw order agent, goal, obstacles; red if speed == 0 else green)
//   crop around the agent          examples/ball_cnn_reinforce.py:130-144 (pad by width / 2 with white, [y-50, y+50) x
//                                  [x-50, x+50) of the padded frame)
//   resize to 40 x 40, ToTensor    examples/ball_cnn_reinforce.py:122-124 (BICUBIC), ball_cnn_ac3.py:253-255 (BILINEAR)
//
// The reference renders through pyglet / OpenGL and resizes through PIL on the host, one environment at a time.  Here
// one block owns one environment and nothing leaves the chip between the state and the 3 x 40 x 40 result:
//   1. the environment's objects (agent, goal, obstacles: a few dozen integers) are staged in shared memory;
//   2. one thread per patch row paints the row's pixels as 3-bit colour codes from precomputed row masks of the three
//      polygons (a polygon centred on an integral point covers a fixed set of pixel centres: the masks are built once
//      on the host with the viewer's geometry), in draw order, so later geometries overwrite earlier ones;
//   3. Pillow's 8-bit two-pass resampling (Resample.c: horizontal pass into a uint8 image, then the vertical pass;
//      22-bit fixed-point coefficients computed exactly as precompute_coeffs / normalize_coeffs_8bpc do) runs on the
//      codes in shared memory: rows no geometry touched take the precomputed all-white result;
//   4. rows of the result go out as float32 (value / 255, ToTensor) or uint8, coalesced along x.
// Bit-exact against oracle/patches.py, whose resize is checked against the installed Pillow (tests/test_patches.py).
// HBM-bound on its output (19.2 KB per environment as float32 against ~300 B of state); no tensor cores.
#pragma once
#include <math.h>
#include <stdint.h>

#include <vector>

#include "ballenv_kernels.cuh"

namespace ballenv {

constexpr int kPatchThreads = 128;
constexpr int kPatchMaxWidth = 128;   // crop size (even)
constexpr int kPatchMaxOut = 64;
constexpr int kPatchObstacleR = 20;   // radius_rand_person, ballenv_env.py:49
constexpr int kPatchAgentR = 5;       // radius_ctrl_person, ballenv_env.py:50
constexpr int kPatchFrame = 500;      // _screen_width / _screen_height, ballenv_env.py:11-12
constexpr int kPatchPrecisionBits = 32 - 8 - 2;

// device tables of one (width, out, filter) combination, built by patch_tables() on the host
struct PatchTables {
  int width, out, ksize;
  const int* bounds;                 // [out][2] first source index, count
  const int* coef;                   // [out][ksize] fixed-point weights
  const unsigned long long* agent;   // [2 * kPatchAgentR] row masks, row 0 = lowest
  const unsigned long long* goal;    // [10]
  const unsigned long long* obstacle;   // [2 * kPatchObstacleR]
};

// ---- host side: the tables --------------------------------------------------------------------------------------------
namespace patch_host {

inline bool in_triangle(double px, double py, const double* a, const double* b, const double* c) {
  const double d1 = (b[0] - a[0]) * (py - a[1]) - (b[1] - a[1]) * (px - a[0]);
  const double d2 = (c[0] - b[0]) * (py - b[1]) - (c[1] - b[1]) * (px - b[0]);
  const double d3 = (a[0] - c[0]) * (py - c[1]) - (a[1] - c[1]) * (px - c[0]);
  return (d1 >= 0 && d2 >= 0 && d3 >= 0) || (d1 <= 0 && d2 <= 0 && d3 <= 0);
}

// row masks of a polygon drawn as a triangle fan from its first vertex (GL_POLYGON / GL_QUADS), centred on an integral
// point: bit ix of row iy <=> the pixel centre (ix - R + 0.5, iy - R + 0.5) is inside
inline void sprite_rows(const std::vector<double>& v, int radius, unsigned long long* rows) {
  const int nv = (int)v.size() / 2;
  for (int iy = 0; iy < 2 * radius; ++iy) {
    unsigned long long m = 0;
    for (int ix = 0; ix < 2 * radius; ++ix) {
      const double px = ix - radius + 0.5, py = iy - radius + 0.5;
      bool in = false;
      for (int k = 1; k + 1 < nv && !in; ++k) in = in_triangle(px, py, &v[0], &v[2 * k], &v[2 * k + 2]);
      if (in) m |= 1ull << ix;
    }
    rows[iy] = m;
  }
}

inline std::vector<double> circle_polygon(double radius, int res = 30) {   // rendering.make_circle
  std::vector<double> v;
  for (int i = 0; i < res; ++i) {
    const double ang = 2 * M_PI * i / res;
    v.push_back(cos(ang) * radius);
    v.push_back(sin(ang) * radius);
  }
  return v;
}

inline double filter_bicubic(double x) {   // Resample.c: bicubic_filter, a = -0.5
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}
inline double filter_bilinear(double x) {
  if (x < 0.0) x = -x;
  return x < 1.0 ? 1.0 - x : 0.0;
}

// Resample.c: precompute_coeffs + normalize_coeffs_8bpc for a [0, in_size) box
inline int coefficients(int in_size, int out_size, bool bicubic, std::vector<int>* bounds, std::vector<int>* kk) {
  const double scale = (double)in_size / out_size;
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = (bicubic ? 2.0 : 1.0) * filterscale;
  const int ksize = (int)ceil(support) * 2 + 1;
  bounds->assign(2 * out_size, 0);
  kk->assign((size_t)out_size * ksize, 0);
  std::vector<double> w(ksize);
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = (xx + 0.5) * scale, ss = 1.0 / filterscale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
      w[x] = bicubic ? filter_bicubic((x + xmin - center + 0.5) * ss) : filter_bilinear((x + xmin - center + 0.5) * ss);
      ww += w[x];
    }
    for (int x = 0; x < xmax; ++x) {
      const double v = ww != 0.0 ? w[x] / ww : w[x];
      (*kk)[(size_t)xx * ksize + x] = v < 0 ? (int)(-0.5 + v * (1 << kPatchPrecisionBits)) : (int)(0.5 + v * (1 << kPatchPrecisionBits));
    }
    (*bounds)[2 * xx] = xmin;
    (*bounds)[2 * xx + 1] = xmax;
  }
  return ksize;
}

}  // namespace patch_host

// dynamic shared memory of one block
__host__ __device__ inline size_t patch_smem_bytes(int width, int out, int ksize, int n_objects) {
  size_t b = ((size_t)width * width + 15) / 16 * 16;   // colour codes
  b += (3 * (size_t)width * out + 15) / 16 * 16;       // horizontal pass, uint8 [3][width][out]
  b += (3 * (size_t)out * out + 15) / 16 * 16;         // result, uint8 [3][out][out]
  b += 4 * ((size_t)out * ksize + 2 * out);            // coefficients, bounds
  b += 4 * (size_t)out;                                // the horizontal result of an all-white row
  b += 4 * 2 * (size_t)width;                          // per row: first / last painted column
  b += 4 * 2 * (size_t)out;                            // per result column: first / last source row that is not white
  b += 4 * 8;                                          // block-wide ranges
  b += 12 * (size_t)n_objects;                         // objects: x, y, kind
  return b;
}

// kind of an object: radius and colour code (bit 0 = R, 1 = G, 2 = B at 255)
constexpr int kObjAgent = 0, kObjGoal = 1, kObjRed = 2, kObjGreen = 3;

// Most of a patch is background: the work follows the painted pixels.  Rows nothing touched, result rows and columns
// whose filter support sees no painted pixel are white by construction (a normalised kernel over 255s gives 255 - the
// precomputed value is used, whatever rounding it carries); the two resampling passes run over the bounding ranges
// only, and the result is assembled in shared memory and streamed out with 128-bit stores.
template <typename T>
__global__ void __launch_bounds__(kPatchThreads) ballenv_patch_kernel(const __grid_constant__ Params p, const PatchTables tb,
                                                                      void* out, int out_u8) {
  extern __shared__ __align__(16) unsigned char sm[];
  const int W = tb.width, O = tb.out, KZ = tb.ksize;
  const int n_obj = 2 + p.cfg.ks + p.cfg.kd;
  unsigned char* const cls = sm;
  unsigned char* const hor = cls + ((size_t)W * W + 15) / 16 * 16;
  unsigned char* const res = hor + (3 * (size_t)W * O + 15) / 16 * 16;
  int* const coef = reinterpret_cast<int*>(res + (3 * (size_t)O * O + 15) / 16 * 16);
  int* const bounds = coef + O * KZ;
  int* const white = bounds + 2 * O;
  int* const rowlo = white + O;
  int* const rowhi = rowlo + W;
  int* const collo = rowhi + W;   // per result column: first / last row of the horizontal pass that is not white
  int* const colhi = collo + O;
  int* const rng = colhi + O;     // 0 / 1: first / last painted row; 2 / 3: first / last painted column; 4..7: result ranges
  int* const obj = rng + 8;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int kWarps = kPatchThreads / 32;
  const long long e = blockIdx.x;
  const int span = W / 2;

  // ---- tables, the environment's objects (coordinates rounded to the nearest integer), white canvases
  for (int i = tid; i < O * KZ; i += kPatchThreads) coef[i] = __ldg(tb.coef + i);
  for (int i = tid; i < 2 * O; i += kPatchThreads) bounds[i] = __ldg(tb.bounds + i);
  for (int k = tid; k < n_obj; k += kPatchThreads) {
    T x, y;
    int kind;
    if (k == 0) {
      x = reinterpret_cast<const T*>(p.agent_x)[e];
      y = reinterpret_cast<const T*>(p.agent_y)[e];
      kind = kObjAgent;
    } else if (k == 1) {
      x = reinterpret_cast<const T*>(p.goal_x)[e];
      y = reinterpret_cast<const T*>(p.goal_y)[e];
      kind = kObjGoal;
    } else if (k < 2 + p.cfg.ks) {
      x = reinterpret_cast<const T*>(p.stat_x)[e * p.stat_stride + (k - 2)];
      y = reinterpret_cast<const T*>(p.stat_y)[e * p.stat_stride + (k - 2)];
      kind = kObjRed;   // obstacles.speed stays 0 (ballenv_env.py:27, 298-301)
    } else {
      const int j = k - 2 - p.cfg.ks;
      x = reinterpret_cast<const T*>(p.dyn_x)[e * p.dyn_stride + j];
      y = reinterpret_cast<const T*>(p.dyn_y)[e * p.dyn_stride + j];
      kind = p.cfg.speed[j] == 0.0 ? kObjRed : kObjGreen;   // :298, :304-306
    }
    obj[3 * k] = (int)rint((double)x);
    obj[3 * k + 1] = (int)rint((double)y);
    obj[3 * k + 2] = kind;
  }
  {
    uint32_t* const c4 = reinterpret_cast<uint32_t*>(cls);
    for (int i = tid; i < (W * W + 3) / 4; i += kPatchThreads) c4[i] = 0x07070707u;   // white; also the padding (:137)
    uint32_t* const r4 = reinterpret_cast<uint32_t*>(res);
    for (int i = tid; i < (3 * O * O + 3) / 4; i += kPatchThreads) r4[i] = 0xffffffffu;
    for (int i = tid; i < O; i += kPatchThreads) {
      collo[i] = W;
      colhi[i] = -1;
    }
    if (tid == 0) {
      rng[0] = W;
      rng[1] = -1;
      rng[2] = W;
      rng[3] = -1;
      rng[4] = O;
      rng[5] = -1;
      rng[6] = O;
      rng[7] = -1;
    }
  }
  __syncthreads();
  // the horizontal pass of a row nothing touched: every source pixel 255 in every channel
  for (int xx = tid; xx < O; xx += kPatchThreads) {
    int s = 1 << (kPatchPrecisionBits - 1);
    for (int x = 0; x < bounds[2 * xx + 1]; ++x) s += 255 * coef[xx * KZ + x];
    white[xx] = min(max(s >> kPatchPrecisionBits, 0), 255);
  }

  // ---- paint: thread i owns patch row i = frame row y (GL rows count upwards; the patch's row 0 is its top); the
  //      geometries in draw order, later ones overwrite
  const int ax = obj[0], ay = obj[1];
  for (int i = tid; i < W; i += kPatchThreads) {
    unsigned char* const row = cls + (size_t)i * W;
    const int y = ay + span - 1 - i;
    int jlo = W, jhi = -1;
    if (y >= 0 && y < kPatchFrame) {
      for (int k = 0; k < n_obj; ++k) {
        const int ox = obj[3 * k], oy = obj[3 * k + 1], kind = obj[3 * k + 2];
        const int r = kind == kObjAgent ? kPatchAgentR : (kind == kObjGoal ? 5 : kPatchObstacleR);
        const int iy = y - oy + r;
        if (iy < 0 || iy >= 2 * r) continue;
        // mask bit ix <-> frame column ox - r + ix <-> patch column j = that - (ax - span)
        const int c0 = ox - r, j0 = c0 - (ax - span);
        const int lo = max(0, max(-j0, -c0)), hi = min(2 * r, min(W - j0, kPatchFrame - c0));
        if (lo >= hi) continue;
        unsigned long long m = __ldg((kind == kObjAgent ? tb.agent : (kind == kObjGoal ? tb.goal : tb.obstacle)) + iy);
        m = (m >> lo) & (hi - lo >= 64 ? ~0ull : ((1ull << (hi - lo)) - 1ull));
        if (m == 0ull) continue;
        const unsigned char colour = kind == kObjRed ? 1 : (kind == kObjGreen ? 2 : 0);
        jlo = min(jlo, j0 + lo + __ffsll((long long)m) - 1);
        jhi = max(jhi, j0 + lo + 63 - __clzll((long long)m));
        while (m != 0ull) {
          const int b = __ffsll((long long)m) - 1;
          m &= m - 1ull;
          row[j0 + lo + b] = colour;
        }
      }
    }
    rowlo[i] = jlo;
    rowhi[i] = jhi;
    if (jhi >= 0) {
      atomicMin(&rng[0], i);
      atomicMax(&rng[1], i);
      atomicMin(&rng[2], jlo);
      atomicMax(&rng[3], jhi);
    }
  }
  __syncthreads();

  if (rng[1] >= 0) {   // (block-uniform) something is painted
    // result rows / columns whose support [first, first + count) meets the painted rows / columns (supports move
    // monotonically with the index), then the source rows those result rows read
    const int i0 = rng[0], i1 = rng[1], j0 = rng[2], j1 = rng[3];
    for (int xx = tid; xx < O; xx += kPatchThreads) {
      const int f = bounds[2 * xx], c = bounds[2 * xx + 1];
      if (f + c > j0 && f <= j1) {
        atomicMin(&rng[4], xx);
        atomicMax(&rng[5], xx);
      }
      if (f + c > i0 && f <= i1) {
        atomicMin(&rng[6], xx);
        atomicMax(&rng[7], xx);
      }
    }
    __syncthreads();
    const int XA = rng[4], XB = rng[5], YA = rng[6], YB = rng[7];
    if (XB >= 0 && YB >= 0) {
      const int I0 = bounds[2 * YA], I1 = bounds[2 * YB] + bounds[2 * YB + 1] - 1;
      // ---- horizontal pass (Resample.c: ImagingResampleHorizontal_8bpc) into uint8 [3][W][O], the needed part; per
      //      result column the rows whose value is not the all-white one are remembered
      for (int i = I0 + warp; i <= I1; i += kWarps) {
        const int lo = rowlo[i], hi = rowhi[i];
        for (int xx = XA + lane; xx <= XB; xx += 32) {
          const int xmin = bounds[2 * xx], cnt = bounds[2 * xx + 1];
          int v0, v1, v2;
          if (hi < 0 || xmin + cnt <= lo || xmin > hi) {
            v0 = v1 = v2 = white[xx];
          } else {
            // colour codes: 7 white, 1 red, 2 green, 0 black - sums of the weights by code, 255 factored out
            int tw = 0, tr = 0, tg = 0;
            const unsigned char* const src = cls + (size_t)i * W + xmin;
            const int* const k = coef + xx * KZ;
            for (int x = 0; x < cnt; ++x) {
              const int c = src[x], kv = k[x];
              tw += c == 7 ? kv : 0;
              tr += c == 1 ? kv : 0;
              tg += c == 2 ? kv : 0;
            }
            const int half = 1 << (kPatchPrecisionBits - 1);
            v0 = min(max((half + 255 * (tw + tr)) >> kPatchPrecisionBits, 0), 255);
            v1 = min(max((half + 255 * (tw + tg)) >> kPatchPrecisionBits, 0), 255);
            v2 = min(max((half + 255 * tw) >> kPatchPrecisionBits, 0), 255);
            if (v0 != white[xx] || v1 != white[xx] || v2 != white[xx]) {
              atomicMin(&collo[xx], i);
              atomicMax(&colhi[xx], i);
            }
          }
          hor[i * O + xx] = (unsigned char)v0;
          hor[W * O + i * O + xx] = (unsigned char)v1;
          hor[2 * W * O + i * O + xx] = (unsigned char)v2;
        }
      }
      __syncthreads();
      // ---- vertical pass (ImagingResampleVertical_8bpc) into the result: only where the support of the result row
      //      meets the column's non-white rows (the rest is the all-white value the result was initialised with)
      for (int yy = YA + warp; yy <= YB; yy += kWarps) {
        const int ymin = bounds[2 * yy], cnt = bounds[2 * yy + 1];
        const int* const k = coef + yy * KZ;
        for (int xx = XA + lane; xx <= XB; xx += 32) {
          if (ymin + cnt <= collo[xx] || ymin > colhi[xx]) continue;
          const unsigned char* const src = hor + (size_t)ymin * O + xx;
          int s0 = 1 << (kPatchPrecisionBits - 1), s1 = s0, s2 = s0;
          for (int y = 0; y < cnt; ++y) {
            const int kv = k[y];
            s0 += (int)src[y * O] * kv;
            s1 += (int)src[W * O + y * O] * kv;
            s2 += (int)src[2 * W * O + y * O] * kv;
          }
          res[yy * O + xx] = (unsigned char)min(max(s0 >> kPatchPrecisionBits, 0), 255);
          res[O * O + yy * O + xx] = (unsigned char)min(max(s1 >> kPatchPrecisionBits, 0), 255);
          res[2 * O * O + yy * O + xx] = (unsigned char)min(max(s2 >> kPatchPrecisionBits, 0), 255);
        }
      }
    }
  }
  __syncthreads();

  // ---- the result rows: [3][O][O], x fastest; float32 = value / 255 (ToTensor), 128-bit streaming stores
  const int total = 3 * O * O;
  if (out_u8) {
    unsigned char* const o = reinterpret_cast<unsigned char*>(out) + (size_t)e * total;
    if ((total & 3) == 0) {
      for (int i = tid; i < total / 4; i += kPatchThreads) reinterpret_cast<uint32_t*>(o)[i] = reinterpret_cast<const uint32_t*>(res)[i];
    } else {
      for (int i = tid; i < total; i += kPatchThreads) o[i] = res[i];
    }
  } else {
    float* const o = reinterpret_cast<float*>(out) + (size_t)e * total;
    if ((total & 3) == 0) {
      for (int i = tid; i < total / 4; i += kPatchThreads) {
        const uint32_t v = reinterpret_cast<const uint32_t*>(res)[i];
        float4 f = make_float4(1.0f, 1.0f, 1.0f, 1.0f);
        if (v != 0xffffffffu) {
          f.x = __fdiv_rn((float)(v & 0xffu), 255.0f);
          f.y = __fdiv_rn((float)((v >> 8) & 0xffu), 255.0f);
          f.z = __fdiv_rn((float)((v >> 16) & 0xffu), 255.0f);
          f.w = __fdiv_rn((float)(v >> 24), 255.0f);
        }
        __stcs(reinterpret_cast<float4*>(o) + i, f);
      }
    } else {
      for (int i = tid; i < total; i += kPatchThreads) __stcs(o + i, __fdiv_rn((float)res[i], 255.0f));
    }
  }
}

}  // namespace ballenv
