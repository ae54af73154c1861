// The batched finish_episode of the actor-critic example as hand-written kernels (sm_100a): loss and parameter gradients
// of Policy(window) over the S = T x N (observation, action, return) triples of a rollout - what
// examples/ball_cnn_ac3.py:222-246 computes with autograd for one episode of one environment.
//
//   policy            examples/ball_cnn_ac3.py:109-146   probs = softmax(action_head(relu(fc1(x)))), v = value_head(relu(fc1(x)))
//   policy loss       :233-238    -log_prob(a) * (R - v.item())        (the advantage is a constant: no gradient through v)
//   value loss        :239        smooth_l1_loss(v, R)
//   loss              :241        sum of both over the steps
// R is the normalised discounted return (:228-232; ballenv_discounted_returns + mean / std on the host side, a2c.py).
//
// Why not autograd: the batched torch update moves the 524 288 x 128 hidden activations of config 5 through HBM a dozen
// times (2.4 ms of a 2.67 ms iteration).  Here nothing but the 29 (or 104) observation floats, the action and the return
// of a sample is read, and nothing but 5 131 partial sums per block is written:
//   phase 1, thread = sample: the observation row becomes bits; fc1 is a sum of the weight columns of the set bits, the
//     heads are accumulated while the hidden units stream by (as in the rollout kernel's policy_action); softmax, log-prob,
//     the sample's loss, d loss / d logits = adv (p - onehot(a)) and d loss / d v = clamp(v - R, -1, 1) go to shared memory;
//   phase 2, thread = hidden unit k, looping over the tile's samples: the lanes of a warp share the sample, so its bits are
//     warp-uniform (no divergence) and the weight reads are conflict-free; h_k is recomputed (a few adds), the gradient
//     flowing into it is a 10-term dot product with the unit's head weights held in registers, and the unit's own sums -
//     d action_head[:, k], d value_head[k], d fc1.bias[k] in registers, d fc1.weight[k, i] for the sample's set bits in the
//     thread's private column of a shared-memory table - need no atomics and no cross-thread reduction;
//   blocks are persistent (a fixed grid walks the tiles), write their partial sums once, and a second kernel adds the
//   partials in double precision in a fixed order: deterministic gradients.
// Checked against torch.autograd on the same batch (tests/test_fused_policy.py).  No tensor cores: the contractions are
// 29 x 128 and 128 x 10 with binary inputs - the work is in the reductions, not in the multiplies.
#pragma once
#include <stdint.h>

namespace ballenv {
namespace a2c {

constexpr int kMaxHidden = 256;
constexpr int kHeads = 10;   // 9 action logits + the value

struct Args {
  int n_in, hidden, tile;     // observation floats, hidden units, samples per tile (= threads per block)
  long long n_samples;
  const float *fc1_w, *fc1_b, *act_w, *act_b, *val_w, *val_b;
  const float* obs;           // [S][n_in]
  const long long* action;    // [S]
  const float* ret;           // [S] returns: normalised already, or raw with ret_stats = {mean, std + eps} on the device
  const float* ret_stats;     // nullptr, or [2]: R = (ret - mean) / (std + eps) is formed here (:231-232)
  const float* pv;            // nullptr, or [S][10]: probabilities and value of the rollout's own forward pass (same weights)
  float* partial;             // [grid][n_params + 1]
};

__host__ __device__ inline int n_params(int n_in, int hidden) { return n_in * hidden + hidden + kHeads * hidden + kHeads; }

__host__ __device__ inline size_t smem_bytes(int n_in, int hidden, int tile, int nw) {
  return 4 * ((size_t)n_in * (hidden + 4) + hidden + (size_t)hidden * 12 + 12 + (size_t)n_in * hidden + (size_t)tile * 12 +
              (size_t)tile * nw);
}

template <int NW>   // 32-bit words of an observation's bits: 1 (WINDOW = 5), 4 (WINDOW = 10)
__global__ void __launch_bounds__(kMaxHidden) grad_kernel(const Args a) {
  extern __shared__ __align__(16) float sm[];
  const int H = a.hidden, HS = H + 4, NB = a.n_in, TS = a.tile, tid = threadIdx.x;
  float* const w1t = sm;                     // fc1.weight transposed [input][hidden + 4]
  float* const b1 = w1t + NB * HS;           // [hidden]
  float* const w2 = b1 + H;                  // heads [hidden][12]: 9 action weights, the value weight, padding
  float* const b2 = w2 + H * 12;             // [12]
  float* const g1 = b2 + 12;                 // d fc1.weight, [input][hidden]: column k belongs to thread k
  float* const dl = g1 + NB * H;             // [tile][12]: d loss / d logits (0..8), d loss / d v (9)
  uint32_t* const bits = reinterpret_cast<uint32_t*>(dl + TS * 12);   // [tile][NW]

  for (int i = tid; i < NB * H; i += TS) {
    const int k = i / NB, in = i - k * NB;
    w1t[in * HS + k] = a.fc1_w[i];
    g1[i] = 0.0f;
  }
  for (int i = tid; i < H; i += TS) b1[i] = a.fc1_b[i];
  for (int i = tid; i < 12 * H; i += TS) {
    const int k = i / 12, j = i - k * 12;
    w2[i] = j < 9 ? a.act_w[j * H + k] : (j == 9 ? a.val_w[k] : 0.0f);
  }
  if (tid < 12) b2[tid] = tid < 9 ? a.act_b[tid] : (tid == 9 ? a.val_b[0] : 0.0f);
  __syncthreads();

  // this thread's hidden unit (phase 2): its head weights and bias in registers, its sums
  const bool unit = tid < H;
  float wk[kHeads], aw[kHeads], ab1 = 0.0f, ab2 = 0.0f, loss = 0.0f;
#pragma unroll
  for (int j = 0; j < kHeads; ++j) {
    wk[j] = unit ? w2[tid * 12 + j] : 0.0f;
    aw[j] = 0.0f;
  }
  const float b1k = unit ? b1[tid] : 0.0f;

  const long long n_tiles = (a.n_samples + TS - 1) / TS;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    // ---- phase 1: thread = sample
    const long long s = tile * TS + tid;
    const bool valid = s < a.n_samples;
    uint32_t x[NW];
#pragma unroll
    for (int w = 0; w < NW; ++w) x[w] = 0u;
    if (valid) {
      const float* const row = a.obs + (size_t)s * NB;
#pragma unroll 1
      for (int b = 0; b < NB; ++b)
        if (row[b] != 0.0f) x[b >> 5] |= 1u << (b & 31);
    }
    float acc[kHeads];
#pragma unroll
    for (int j = 0; j < kHeads; ++j) acc[j] = b2[j];
    const bool stored = a.pv != nullptr;   // (block-uniform) the rollout kernel kept its softmax and value: no forward pass here
#pragma unroll 1
    for (int k0 = 0; k0 < (stored ? 0 : H); k0 += 4) {
      float4 h = *reinterpret_cast<const float4*>(b1 + k0);
#pragma unroll
      for (int w = 0; w < NW; ++w) {
        uint32_t m = x[w];
        while (m != 0u) {
          const int i = __ffs((int)m) - 1 + 32 * w;
          m &= m - 1u;
          const float4 c = *reinterpret_cast<const float4*>(w1t + i * HS + k0);
          h.x += c.x;
          h.y += c.y;
          h.z += c.z;
          h.w += c.w;
        }
      }
      const float hv[4] = {fmaxf(h.x, 0.0f), fmaxf(h.y, 0.0f), fmaxf(h.z, 0.0f), fmaxf(h.w, 0.0f)};
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const float4* const r = reinterpret_cast<const float4*>(w2 + (k0 + kk) * 12);
        const float4 p0 = r[0], p1 = r[1], p2 = r[2];
        acc[0] = fmaf(hv[kk], p0.x, acc[0]);
        acc[1] = fmaf(hv[kk], p0.y, acc[1]);
        acc[2] = fmaf(hv[kk], p0.z, acc[2]);
        acc[3] = fmaf(hv[kk], p0.w, acc[3]);
        acc[4] = fmaf(hv[kk], p1.x, acc[4]);
        acc[5] = fmaf(hv[kk], p1.y, acc[5]);
        acc[6] = fmaf(hv[kk], p1.z, acc[6]);
        acc[7] = fmaf(hv[kk], p1.w, acc[7]);
        acc[8] = fmaf(hv[kk], p2.x, acc[8]);
        acc[9] = fmaf(hv[kk], p2.y, acc[9]);
      }
    }
    {
      float mx = 0.0f, e[9], sum = 1.0f;
      if (stored) {   // the probabilities and the value the rollout kernel computed with these weights
        const float* const q = a.pv + (size_t)(valid ? s : 0) * 10;
#pragma unroll
        for (int j = 0; j < 9; ++j) e[j] = q[j];
        acc[9] = q[9];
      } else {
        mx = acc[0];
#pragma unroll
        for (int j = 1; j < 9; ++j) mx = fmaxf(mx, acc[j]);
        sum = 0.0f;
#pragma unroll
        for (int j = 0; j < 9; ++j) {
          e[j] = expf(acc[j] - mx);
          sum += e[j];
        }
      }
      const int act = valid ? (int)a.action[s] : 0;
      float R = valid ? a.ret[s] : 0.0f;
      if (a.ret_stats != nullptr) R = valid ? __fdiv_rn(__fsub_rn(R, a.ret_stats[0]), a.ret_stats[1]) : 0.0f;
      const float v = acc[9];
      const float adv = R - v;                                  // :234 reward = r - value.item()
      float la = acc[0], pa = e[0];
#pragma unroll
      for (int j = 1; j < 9; ++j) {
        la = act == j ? acc[j] : la;
        pa = act == j ? e[j] : pa;
      }
      const float logp = stored ? logf(fmaxf(pa, 1e-38f)) : (la - mx) - logf(sum);   // log(probs[a]), :218
      const float d = v - R, ad = fabsf(d);
      if (valid) loss += -logp * adv + (ad < 1.0f ? 0.5f * d * d : ad - 0.5f);   // :237-239
      const float inv = 1.0f / sum;
      float* const o = dl + tid * 12;
#pragma unroll
      for (int j = 0; j < 9; ++j) o[j] = valid ? adv * (e[j] * inv - (act == j ? 1.0f : 0.0f)) : 0.0f;
      o[9] = valid ? fminf(fmaxf(d, -1.0f), 1.0f) : 0.0f;
      o[10] = o[11] = 0.0f;
#pragma unroll
      for (int w = 0; w < NW; ++w) bits[tid * NW + w] = x[w];
    }
    __syncthreads();
    // ---- phase 2: thread = hidden unit, over the tile's samples (warp-uniform bits)
    if (unit) {
#pragma unroll 2
      for (int s2 = 0; s2 < TS; ++s2) {
        const float4* const dr = reinterpret_cast<const float4*>(dl + s2 * 12);
        const float4 d0 = dr[0], d1 = dr[1], d2 = dr[2];
        float h = b1k;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
          uint32_t m = bits[s2 * NW + w];
          while (m != 0u) {
            const int i = __ffs((int)m) - 1 + 32 * w;
            m &= m - 1u;
            h += w1t[i * HS + tid];
          }
        }
        const float dv[kHeads] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w, d2.x, d2.y};
        float gsum = 0.0f;
        const float hr = fmaxf(h, 0.0f);
#pragma unroll
        for (int j = 0; j < kHeads; ++j) {
          gsum = fmaf(wk[j], dv[j], gsum);
          aw[j] = fmaf(dv[j], hr, aw[j]);
        }
        const float dh = h > 0.0f ? gsum : 0.0f;
        ab1 += dh;
        if (dh != 0.0f) {
#pragma unroll
          for (int w = 0; w < NW; ++w) {
            uint32_t m = bits[s2 * NW + w];
            while (m != 0u) {
              const int i = __ffs((int)m) - 1 + 32 * w;
              m &= m - 1u;
              g1[i * H + tid] += dh;
            }
          }
        }
      }
    }
    if (tid < kHeads) {   // d bias of the heads: the column sums of the tile's table
      for (int s2 = 0; s2 < TS; ++s2) ab2 += dl[s2 * 12 + tid];
    }
    __syncthreads();
  }

  // ---- this block's partial sums: [d fc1.weight [hidden][input]] [d fc1.bias] [d heads [10][hidden]] [d head biases] [loss]
  const int P = n_params(NB, H);
  float* const out = a.partial + (size_t)blockIdx.x * (P + 1);
  for (int i = tid; i < NB * H; i += TS) {
    const int k = i / NB, in = i - k * NB;
    out[i] = g1[in * H + k];
  }
  if (unit) {
    out[NB * H + tid] = ab1;
#pragma unroll
    for (int j = 0; j < kHeads; ++j) out[NB * H + H + j * H + tid] = aw[j];
  }
  if (tid < kHeads) out[NB * H + H + kHeads * H + tid] = ab2;
  // the loss: block sum through shared memory (dl is free now)
  dl[tid] = loss;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (int i = 0; i < TS; ++i) t += (double)dl[i];
    out[P] = (float)t;
  }
}

struct Outs {
  float *fc1_w, *fc1_b, *act_w, *act_b, *val_w, *val_b, *loss;
};

// sums of the blocks' partials, in double, in a fixed order: 32 parameters per block of 256 threads, eight threads per
// parameter each adding every eighth block's partial, then the eight in order (deterministic; 592 partials per
// parameter summed by one thread were a 47 us chain of dependent loads)
__global__ void __launch_bounds__(256) reduce_kernel(const float* __restrict__ partial, int n_blocks, int n_in, int hidden, Outs o) {
  __shared__ double part[8][32];
  const int P = n_params(n_in, hidden);
  const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + lane;
  double t = 0.0;
  if (i <= P)
    for (int b = slice; b < n_blocks; b += 8) t += (double)partial[(size_t)b * (P + 1) + i];
  part[slice][lane] = t;
  __syncthreads();
  if (slice != 0 || i > P) return;
  for (int q = 1; q < 8; ++q) t += part[q][lane];
  const float v = (float)t;
  const int H = hidden, NB = n_in;
  if (i < NB * H) o.fc1_w[i] = v;
  else if (i < NB * H + H) o.fc1_b[i - NB * H] = v;
  else if (i < NB * H + H + 9 * H) o.act_w[i - NB * H - H] = v;
  else if (i < NB * H + H + 10 * H) o.val_w[i - NB * H - H - 9 * H] = v;
  else if (i < P - 1) o.act_b[i - (NB * H + H + 10 * H)] = v;
  else if (i == P - 1) o.val_b[0] = v;
  else o.loss[0] = v;
}

}  // namespace a2c
}  // namespace ballenv
