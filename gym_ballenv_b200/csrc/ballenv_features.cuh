// 20-float social-navigation feature vector of the pygame ruleset (sm_100a): createBoard.sensor_readings.
//
//   featureExtractor.featureExtractor   featureExtractor.py:247-265  (called by ballenv_pygame.py:512, 674)
//     [0]      min(floor(dist_goal / 5), 5)                                  calcDistanceFromGoal   :132-144
//     [1:5]    goal direction one-hot against the +y axis (45 / 135 degree sectors, left / right)   :146-166
//     [5:8]    obstacles with surface distance < 101, < 230, < 1000         densityFeatures        :91-112
//     [8:17]   3 x 3 histogram orientation bin x relative-speed bin         speedOrientationFeatures :115-130, 61-86
//     [17:20]  social force per orientation bin, f = exp(-d / 10) * d * (2 - 0.5 (1 + cos psi)), if f > 1   :170-193
//   surface distance d = |obstacle - agent| - agent_rad - obstacle.rad      calcDistance           :36-40
//
// The reference always passes agent velocity (0, 0) (ballenv_pygame.py:338-339, 674) and its obstacles never move
// (vel 0, :24-38), and Obstacle(self.rad_static_obstacles) passes the radius as the id, so obstacle.rad is always
// 20 (:492, :33-36).  Velocities are kernel parameters here (0 today) so the formulas are the reference's in full.
//
// One thread per environment; a secondary observe mode (80 bytes out per environment), not the step hot path.
#pragma once
#include <stdint.h>

#include "ballenv_kernels.cuh"

namespace ballenv {

constexpr double kFeatureObstacleRad = 20.0;   // Obstacle.rad default, ballenv_pygame.py:33-36

template <typename T>
__device__ __forceinline__ T angle_between(T ax, T ay, T bx, T by) {
  // featureExtractor.py:43-56: unit vectors (zero vectors stay zero), arccos(clip(dot, -1, 1))
  const T na = sqrt(ax * ax + ay * ay), nb = sqrt(bx * bx + by * by);
  if (na > (T)0) { ax /= na; ay /= na; }
  if (nb > (T)0) { bx /= nb; by /= nb; }
  T d = ax * bx + ay * by;
  d = d < (T)-1 ? (T)-1 : (d > (T)1 ? (T)1 : d);
  return acos(d);
}

template <typename T>
__global__ void __launch_bounds__(128) ballenv_features_kernel(const __grid_constant__ Params p, float* __restrict__ out,
                                                               double agent_rad, double agent_vx, double agent_vy,
                                                               double obst_vx, double obst_vy) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= p.n) return;
  const T kPi = (T)3.14159265358979323846;
  const T ax = reinterpret_cast<const T*>(p.agent_x)[e], ay = reinterpret_cast<const T*>(p.agent_y)[e];
  const T gx = reinterpret_cast<const T*>(p.goal_x)[e], gy = reinterpret_cast<const T*>(p.goal_y)[e];
  float f[20];
#pragma unroll
  for (int i = 0; i < 20; ++i) f[i] = 0.0f;

  // [0] distance-from-goal bin (:132-144)
  const T dg = floor(hypot(ax - gx, ay - gy) / (T)5);
  f[0] = dg > (T)5 ? 5.0f : (float)dg;
  // [1:5] goal direction (:146-166)
  {
    const T xi = gx - ax, yi = gy - ay;
    const T ang = angle_between<T>((T)0, (T)1, xi, yi);
    if (ang < kPi / 4) f[1] = 1.0f;
    else if (ang > kPi / 4 && ang < kPi * 3 / 4) f[xi > (T)0 ? 2 : 4] = 1.0f;
    else f[3] = 1.0f;
  }
  // obstacles: density, orientation x speed histogram, social forces
  const T rvx = (T)(obst_vx - agent_vx), rvy = (T)(obst_vy - agent_vy);
  const T relvel = sqrt(rvx * rvx + rvy * rvy);
  const int speed_bin = relvel < (T)0.015 ? 0 : (relvel < (T)0.025 ? 1 : 2);      // :66-72
  const int K = p.cfg.ks + p.cfg.kd;
  for (int k = 0; k < K; ++k) {
    T ox, oy;
    if (k < p.cfg.ks) {
      ox = reinterpret_cast<const T*>(p.stat_x)[e * p.stat_stride + k];
      oy = reinterpret_cast<const T*>(p.stat_y)[e * p.stat_stride + k];
    } else {
      ox = reinterpret_cast<const T*>(p.dyn_x)[e * p.dyn_stride + (k - p.cfg.ks)];
      oy = reinterpret_cast<const T*>(p.dyn_y)[e * p.dyn_stride + (k - p.cfg.ks)];
    }
    const T dx = ox - ax, dy = oy - ay;
    const T d = sqrt(dx * dx + dy * dy) - (T)agent_rad - (T)kFeatureObstacleRad;   // :36-40
    if (d < (T)1000) f[7] += 1.0f;                                                   // :103-108
    if (d < (T)230) f[6] += 1.0f;
    if (d < (T)101) f[5] += 1.0f;
    const T psi = angle_between<T>(dx, dy, rvx, rvy);                                // :74
    const int obin = psi < kPi / 4 ? 0 : ((psi > kPi / 4 && psi < kPi * 3 / 4) ? 1 : 2);   // :76-83
    f[8 + obin * 3 + speed_bin] += 1.0f;                                             // :122-124
    const T fsoc = exp(-d / (T)10) * d * ((T)2 + (T)0.5 * ((T)1 - (T)2) * ((T)1 + cos(psi)));   // :186-189
    if (fsoc > (T)1) f[17 + obin] += (float)fsoc;                                    // :191-192
  }
  float4* o = reinterpret_cast<float4*>(out + e * 20);   // 80-byte rows: 16-byte aligned
#pragma unroll
  for (int i = 0; i < 5; ++i) o[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
}

// Legacy 29-float observation of the REINFORCE / imitation / supervised scripts: prep_state2 + block_to_arrpos of
// examples/ball_env_reinforce.py:130-172 (the same function in ball_env_imitate.py and test_model.py; the shipped
// path logs hold these vectors).  [0:4] goal-quadrant bits (:139-150); [4:29] a 5 x 5 grid of obstacle COUNTS in
// 20-pixel blocks around the agent, row = y block, the agent's own cell (index 4 + 12) starting at 1 (:138).
//   dx = agent_x - obstacle_x; block_x = sign(dx) * (dx - 10) // 20, likewise y, unless dx == 0 or dy == 0 -> (0, 0)
//   (:153-158: dx > 0 gives floor((dx - 10) / 20), dx < 0 gives floor((|dx| + 10) / 20));
//   counted if |block_x| < 3 and |block_y| < 3 at index 4 + 12 + 5 * block_y + block_x (:160-164, :169-172).
// One thread per environment, obstacles in list order (static, then dynamic); a secondary observe mode.
template <typename T>
__global__ void __launch_bounds__(128) ballenv_blocks_kernel(const __grid_constant__ Params p, float* __restrict__ out) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= p.n) return;
  const double ax = (double)reinterpret_cast<const T*>(p.agent_x)[e], ay = (double)reinterpret_cast<const T*>(p.agent_y)[e];
  const double gx = (double)reinterpret_cast<const T*>(p.goal_x)[e], gy = (double)reinterpret_cast<const T*>(p.goal_y)[e];
  float f[29];
#pragma unroll
  for (int i = 0; i < 29; ++i) f[i] = 0.0f;
  f[12 + 4] = 1.0f;
  const int q = goal_quadrant_bit(gx - ax < 0.0, gy - ay < 0.0);
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (i == q) f[i] = 1.0f;
  const int K = p.cfg.ks + p.cfg.kd;
  for (int k = 0; k < K; ++k) {
    double ox, oy;
    if (k < p.cfg.ks) {
      ox = (double)reinterpret_cast<const T*>(p.stat_x)[e * p.stat_stride + k];
      oy = (double)reinterpret_cast<const T*>(p.stat_y)[e * p.stat_stride + k];
    } else {
      ox = (double)reinterpret_cast<const T*>(p.dyn_x)[e * p.dyn_stride + (k - p.cfg.ks)];
      oy = (double)reinterpret_cast<const T*>(p.dyn_y)[e * p.dyn_stride + (k - p.cfg.ks)];
    }
    const double xd = ax - ox, yd = ay - oy;
    double xb = 0.0, yb = 0.0;
    if (xd != 0.0 && yd != 0.0) {
      xb = floor((xd > 0.0 ? 1.0 : -1.0) * (xd - 10.0) / 20.0);
      yb = floor((yd > 0.0 ? 1.0 : -1.0) * (yd - 10.0) / 20.0);
    }
    if (fabs(xb) < 3.0 && fabs(yb) < 3.0) {
      const int pos = 4 + 12 + 5 * (int)yb + (int)xb;
#pragma unroll
      for (int i = 4; i < 29; ++i)   // registers, not a local array
        if (i == pos) f[i] += 1.0f;
    }
  }
  float* o = out + e * 29;
#pragma unroll
  for (int i = 0; i < 29; ++i) o[i] = f[i];
}

}  // namespace ballenv
