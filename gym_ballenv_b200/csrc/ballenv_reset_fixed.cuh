// createBoard.resetFixedstate (ballenv_pygame.py:589-624) for the environments of a handle (sm_100a): a new episode
// with the goal at a fixed point ((145, 120) in the reference) and the OBSTACLES KEPT:
//   repeat
//     agent = (ranf * 100, ranf * 100)                       :601-602, generate_randomval :454-457
//     dist  = distance(goal, agent)  -> state[2], old_dist    :605-606  (of this first draw, kept if redrawn)
//     while distance(goal, agent) < 50: redraw the agent      :607-612
//   until the agent touches no obstacle                        :613-616 (calc_reward: check_overlap :381-387, :683-688;
//                                                              its goal test cannot fire: the goal is >= 49 away)
//   total_reward_accumulated = 0, total_distance = distance(agent, goal)   :621-622
// One thread per environment, fp64 (the pygame ruleset always is); a rare, host-driven call, not the step hot path.
// Draws: Philox block (global env id, episode, kResetFixedAgent << 28 | outer << 12 | inner, reset stream); the block's
// four words are the two ranf of one agent draw (inner = 0: the first draw of an outer attempt, inner >= 1: redraws).
#pragma once
#include <stdint.h>

#include "ballenv_kernels.cuh"

namespace ballenv {

constexpr uint32_t kResetFixedAgent = 4;
constexpr int kMaxFixedOuter = 65536, kMaxFixedInner = 4096;

__global__ void __launch_bounds__(128) ballenv_reset_fixed_kernel(const __grid_constant__ Params p, const uint8_t* __restrict__ mask,
                                                                  double goal_x, double goal_y) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= p.n || (mask != nullptr && mask[e] == 0)) return;
  const DevConfig& cfg = p.cfg;
  const uint32_t g = p.g0 + (uint32_t)e;
  const uint32_t episode = p.episode[e] + 1;
  const double* sx = reinterpret_cast<const double*>(p.stat_x) + e * p.stat_stride;
  const double* sy = reinterpret_cast<const double*>(p.stat_y) + e * p.stat_stride;
  double ax = 0.0, ay = 0.0, dist = 0.0;
  for (int outer = 0;; ++outer) {
    uint4 w = philox4x32_10(g, episode, (kResetFixedAgent << 28) | ((uint32_t)outer << 12), kStreamReset, p.k0, p.k1);
    ax = 0.0 + ranf_from_words(w.x, w.y) * (cfg.world_w - 0.0);
    ay = 0.0 + ranf_from_words(w.z, w.w) * (cfg.world_h - 0.0);
    dist = dist64(goal_x, goal_y, ax, ay);
    for (int inner = 1; dist64(goal_x, goal_y, ax, ay) < 50.0; ++inner) {
      w = philox4x32_10(g, episode, (kResetFixedAgent << 28) | ((uint32_t)outer << 12) | (uint32_t)inner, kStreamReset, p.k0,
                        p.k1);
      ax = 0.0 + ranf_from_words(w.x, w.y) * (cfg.world_w - 0.0);
      ay = 0.0 + ranf_from_words(w.z, w.w) * (cfg.world_h - 0.0);
      if (inner >= kMaxFixedInner - 1) {
        atomicOr(p.errors, (uint32_t)BALLENV_DEVERR_RESET_STUCK);
        break;
      }
    }
    bool hit = false;
    for (int k = 0; k < cfg.ks; ++k)
      hit = hit || !(dist64(ax, ay, sx[k], sy[k]) > cfg.radius_sum);                 // check_overlap, thresh = 0
    if (!hit) break;
    if (outer >= kMaxFixedOuter - 1) {
      atomicOr(p.errors, (uint32_t)BALLENV_DEVERR_RESET_STUCK);
      break;
    }
  }
  p.episode[e] = episode;
  reinterpret_cast<double*>(p.agent_x)[e] = ax;
  reinterpret_cast<double*>(p.agent_y)[e] = ay;
  reinterpret_cast<double*>(p.goal_x)[e] = goal_x;
  reinterpret_cast<double*>(p.goal_y)[e] = goal_y;
  p.dist[e] = dist;
  p.total[e] = dist64(ax, ay, goal_x, goal_y);
  p.acc[e] = 0.0;
  p.ep_len[e] = 0;
  p.flags[e] = 0;
}

}  // namespace ballenv
