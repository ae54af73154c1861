"""CPU oracle: a plain-Python restatement of the reference's step()/reset()/
window-observation hot path.

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
CPU-baseline / ``--impl reference`` legs of ``bench.py`` import this; the
product (``gym_ballenv_b200``) never does and has no CPU fallback.

Parity pin: ``tests/test_oracle_vs_golden.py`` checks this file against the
fixtures in ``tests/golden/`` that ``oracle/gen_golden.py`` recorded by running
the reference's own, unedited code (``oracle/ref_shim.py``).  The 1000-step
gym ``TimeLimit`` is third-party (gym==0.10.9, not vendored in the reference):
that single row is "parity unpinned" and restated from its documented
behaviour (done once elapsed_steps >= max_episode_steps).

It is deliberately scalar Python with ``math.sqrt``/``math.pow`` like the
reference, so that timing it is a fair stand-in for the reference's own CPU
cost (``cpu_baseline.kind == "port"``).  ``oracle/ballenv_oracle.c`` is the same
algorithm in C (checked against this file by tests/test_oracle_c.py) for parity at full sizes.

Reference citations (relative to /root/reference):
  step            gym_ballenv/envs/ballenv_env.py:232-289
  reward / hit    gym_ballenv/envs/ballenv_env.py:200-229, 179-191
  obstacle motion gym_ballenv/envs/ballenv_env.py:323-353
  reset           gym_ballenv/envs/ballenv_env.py:113-167, 19-33, 193-197
  window obs      examples/ball_cnn_ac3.py:330-352 (goal quadrant), 384-412 (W x W raster)
  pygame ruleset  ballenv_pygame.py:650-706 (step, calc_reward), 460-513 (reset), 381-387
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Sequence, Tuple

from . import draws as D

RULESET_GYM = 0
RULESET_PYGAME = 1

# agent action table of the training loops (examples/ball_cnn_ac3.py:530)
AGENT_MOVES = [(1, 1), (1, -1), (1, 0), (0, 1), (0, -1), (0, 0), (-1, 1), (-1, 0), (-1, -1)]
# obstacle action table (ballenv_env.py:324): (-1,-1) twice, no (-1,0)
OBSTACLE_MOVES = [(1, 1), (1, -1), (1, 0), (0, 1), (0, -1), (0, 0), (-1, 1), (-1, -1), (-1, -1)]


@dataclass
class OracleConfig:
    ruleset: int = RULESET_GYM
    window: int = 5
    n_static: int = 13
    n_dynamic: int = 5
    speeds: Sequence[float] = (1, 1, 1, 1, 1)
    goals: Sequence[Tuple[int, int]] = ((12, 122), (123, 93), (87, 150), (430, 440), (230, 11))
    change_step: int = 50
    rd_th_obs: int = 60
    static_penalty: float = 1          # static_penalty[1]
    dynamic_penalty: float = 8000      # dynamic_penalty[1]
    max_episode_steps: int = 1000      # gym_ballenv/__init__.py:7 ; 0 = no limit
    auto_reset: bool = True
    # pygame ruleset ctor arguments (ballenv_pygame.py:316)
    agent_radius: float = 10
    static_obstacle_radius: float = 10

    @property
    def world(self):
        return (500, 500) if self.ruleset == RULESET_GYM else (100, 100)


def _dist(p, q):
    # calculate_distance, ballenv_env.py:179-183 / ballenv_pygame.py:375-379
    return math.sqrt(math.pow(p[0] - q[0], 2) + math.pow(p[1] - q[1], 2))


def goal_quadrant(agent, goal):
    """prep_state2 (examples/ball_cnn_ac3.py:330-352): index of the hot bit."""
    dx = goal[0] - agent[0]
    dy = goal[1] - agent[1]
    if dx >= 0 and dy >= 0:
        return 1
    if dx < 0 and dy >= 0:
        return 0
    if dx < 0 and dy < 0:
        return 3
    return 2


def window_rows(agent, obstacles, window, radius_sum=25, step=(1, 1)):
    """W x W occupancy raster of prep_state4 (examples/ball_cnn_ac3.py:384-412)
    as a list of W row bitmasks (bit c = column c).

    Column c samples x = agent_x - step_x*int(W/2) + step_x*c.  The row
    coordinate is advanced *after* the column loop with the current r
    (:409), so rows 0 and 1 both sample y = agent_y - step_y*int(W/2) and row
    r >= 1 samples start_y + step_y*(r-1).
    """
    h = int(window / 2)
    sx = agent[0] - step[0] * h
    sy = agent[1] - step[1] * h
    rows = []
    y = sy
    for r in range(window):
        bits = 0
        for c in range(window):
            x = sx + step[0] * c
            for o in obstacles:
                if not (_dist((x, y), o) > radius_sum):
                    bits |= 1 << c
                    break
        rows.append(bits)
        y = sy + step[1] * r
    return rows


def window_obs(agent, goal, obstacles, window, radius_sum=25):
    """prep_state4 as a flat list of 4 + W*W floats (0.0 / 1.0)."""
    out = [0.0] * (4 + window * window)
    out[goal_quadrant(agent, goal)] = 1.0
    for r, bits in enumerate(window_rows(agent, obstacles, window, radius_sum)):
        for c in range(window):
            if bits >> c & 1:
                out[4 + r * window + c] = 1.0
    return out


class OracleEnv:
    """One environment; draws come from an addressed source (oracle.draws)."""

    def __init__(self, cfg: OracleConfig, source, g: int = 0):
        self.cfg = cfg
        self.src = source
        self.g = g
        self.episode = -1
        self.tick = 0
        self.ep_len = 0
        self.agent = (0, 0)
        self.goal = (0, 0)
        self.dist = 0.0
        self.total_distance = 1.0
        self.acc = 0.0
        self.obst: List[List[float]] = []
        self.goal_idx: List[int] = []
        self.counter: List[int] = []
        # flags of the last step
        self.goal_flag = False
        self.hit = False
        self.hit_index = -1
        self.truncated = False
        if cfg.ruleset == RULESET_GYM:
            self.radius_sum = 20 + 5                       # ballenv_env.py:49-50,188
            self.goal_threshold = 10                       # :64
        else:
            self.radius_sum = cfg.static_obstacle_radius + cfg.agent_radius   # ballenv_pygame.py:384
            self.goal_threshold = 15                       # :345

    # ------------------------------------------------------------------ reset
    def reset(self):
        self.episode += 1
        self.ep_len = 0
        if self.cfg.ruleset == RULESET_GYM:
            self._reset_gym()
        else:
            self._reset_pygame()
        return self.state()

    def reset_fixed(self, goal=(145, 120)):
        """createBoard.resetFixedstate (ballenv_pygame.py:589-624): goal fixed, obstacles kept, the agent redrawn until it
        is >= 50 from the goal (:607-612) and touches no obstacle (:613-616, calc_reward :680-688)."""
        assert self.cfg.ruleset != RULESET_GYM
        src, g = self.src, self.g
        self.episode += 1
        self.ep_len = 0
        ep = self.episode
        outer = 0
        while True:
            w = src.reset_words(g, ep, D.RK_FIXED_AGENT, item=outer, attempt=0, count=4)
            agent = (0 + D.ranf_from_words(w[0], w[1]) * (100 - 0),        # :601-602, 454-457
                     0 + D.ranf_from_words(w[2], w[3]) * (100 - 0))
            dist = _dist(goal, agent)                                      # :605 (kept even if redrawn)
            inner = 1
            while _dist(goal, agent) < 50:                                 # :607-612
                w = src.reset_words(g, ep, D.RK_FIXED_AGENT, item=outer, attempt=inner, count=4)
                agent = (0 + D.ranf_from_words(w[0], w[1]) * (100 - 0),
                         0 + D.ranf_from_words(w[2], w[3]) * (100 - 0))
                inner += 1
            if not any(not (_dist(agent, o) > self.radius_sum) for o in self.obst):   # :613-616 -> check_overlap :381-387
                break
            outer += 1
        self.agent, self.goal, self.dist = agent, tuple(goal), dist
        self.acc = 0.0                                                     # :621
        self.total_distance = _dist(agent, goal)                           # :622
        return self.state()

    def _reset_gym(self):
        cfg, src, g, ep = self.cfg, self.src, self.g, self.episode
        w = src.reset_words(g, ep, D.RK_HEAD, count=4)
        goal = (D.mulhi(w[0], 500), 480 + D.mulhi(w[1], 20))     # :115-116
        agent = (D.mulhi(w[2], 500), D.mulhi(w[3], 10))          # :117-118
        dist = _dist(goal, agent)                                # :119
        attempt = 0
        while _dist(goal, agent) < 50:                           # :121-126 (dead for this geometry)
            w = src.reset_words(g, ep, D.RK_AGENT_REDRAW, attempt=attempt, count=2)
            agent = (D.mulhi(w[0], 500), D.mulhi(w[1], 10))
            attempt += 1
        self.agent, self.goal, self.dist = agent, goal, dist
        self.acc = 0.0
        self.obst, self.goal_idx, self.counter = [], [], []
        for i in range(cfg.n_static):                            # :131-149
            attempt = 0
            while True:
                w = src.reset_words(g, ep, D.RK_STATIC, i, attempt)
                x, y = D.mulhi(w[0], 500), 20 + D.mulhi(w[1], 460)   # :24-25
                attempt += 1
                if not self._rect(x, y, agent) and not self._rect(x, y, goal):
                    self.obst.append([x, y])
                    break
        for j in range(cfg.n_dynamic):                           # :153-164
            w = src.reset_words(g, ep, D.RK_DYNAMIC, j)
            self.obst.append([D.mulhi(w[0], 500), 20 + D.mulhi(w[1], 460)])
            self.goal_idx.append(j)
            self.counter.append(0)
        self.total_distance = _dist(agent, goal)                 # :166

    @staticmethod
    def _rect(x, y, p, rad=20, agent_rad=5):
        # check_overlap_rect, ballenv_env.py:193-197
        return abs(x - p[0]) < (rad + agent_rad) and abs(y - p[1]) < (rad / 2 + agent_rad)

    def _reset_pygame(self):
        cfg, src, g, ep = self.cfg, self.src, self.g, self.episode
        w = src.reset_words(g, ep, D.RK_HEAD, item=0, count=4)
        goal = (0 + D.ranf_from_words(w[0], w[1]) * (100 - 0),        # ballenv_pygame.py:468-469,454-457
                0 + D.ranf_from_words(w[2], w[3]) * (100 - 0))
        w = src.reset_words(g, ep, D.RK_HEAD, item=1, count=4)
        agent = (0 + D.ranf_from_words(w[0], w[1]) * (100 - 0),       # :472-473
                 0 + D.ranf_from_words(w[2], w[3]) * (100 - 0))
        dist = _dist(goal, agent)                                     # :474 (kept even if redrawn, :482)
        attempt = 0
        while _dist(goal, agent) < 50:                                # :476-481
            w = src.reset_words(g, ep, D.RK_AGENT_REDRAW, attempt=attempt, count=4)
            agent = (0 + D.ranf_from_words(w[0], w[1]) * (100 - 0),
                     0 + D.ranf_from_words(w[2], w[3]) * (100 - 0))
            attempt += 1
        self.agent, self.goal, self.dist = agent, goal, dist
        self.acc = 0.0
        self.obst, self.goal_idx, self.counter = [], [], []
        for i in range(cfg.n_static):                                 # :489-498
            attempt = 0
            while True:
                w = src.reset_words(g, ep, D.RK_STATIC, i, attempt)
                x, y = D.mulhi(w[0], 100), D.mulhi(w[1], 100)         # :27,32
                attempt += 1
                if (_dist((x, y), agent) - 15 > self.radius_sum) and (_dist((x, y), goal) - 5 > self.radius_sum):
                    self.obst.append([x, y])
                    break
        self.total_distance = _dist(agent, goal)                      # :511

    # ------------------------------------------------------------------ step
    def _move_obstacle(self, j):
        """move_obstacles, ballenv_env.py:323-353."""
        cfg = self.cfg
        o = self.obst[cfg.n_static + j]
        s = cfg.speeds[j]
        if self.counter[j] < cfg.change_step:
            gx, gy = cfg.goals[self.goal_idx[j]]
            tx, ty = gx - o[0], gy - o[1]
            w1 = self.src.step_word(self.g, self.tick, j)
            if tx != 0 and ty != 0:
                if D.mulhi(w1, 100) < cfg.rd_th_obs:
                    o[0] += (tx / abs(tx)) * s
                    o[1] += (ty / abs(ty)) * s
                else:
                    m = OBSTACLE_MOVES[D.mulhi(self.src.step_word2(self.g, self.tick, j, w1, 100), 9)]
                    o[0] += m[0] * s
                    o[1] += m[1] * s
            else:
                m = OBSTACLE_MOVES[D.mulhi(w1, 9)]
                o[0] += m[0] * s
                o[1] += m[1] * s
            self.counter[j] += 1
        else:
            cur = cfg.goals[self.goal_idx[j]]
            others = [k for k, gl in enumerate(cfg.goals) if tuple(gl) != tuple(cur)]
            w1 = self.src.step_word(self.g, self.tick, j)
            self.goal_idx[j] = others[D.mulhi(w1, len(others))]
            self.counter[j] = 0

    def step(self, action):
        """-> (reward, done).  No time limit, no auto-reset (see OracleVec)."""
        cfg = self.cfg
        w, h = cfg.world
        if cfg.ruleset == RULESET_GYM:
            old = self.dist                                   # :236
            nx = self.agent[0] + 1 * action[0]                # :247-250 (speed 1)
            ny = self.agent[1] + 1 * action[1]
            if nx < 0:
                nx = 0
            if ny < 0:
                ny = 0
            if nx > w:
                nx = w
            if ny > h:
                ny = h
            for j in range(cfg.n_dynamic):                    # :262-264
                self._move_obstacle(j)
            self.agent = (nx, ny)
            self.dist = _dist(self.goal, self.agent)          # :268
            self.goal_flag = self.dist < self.goal_threshold  # :276
            reward = -0 + (old - self.dist) / self.total_distance     # :205-206
            self.hit, self.hit_index = False, -1
            for k, o in enumerate(self.obst):                 # :208-224
                if not (_dist(self.agent, o) > self.radius_sum):
                    self.hit, self.hit_index = True, k
                    reward -= cfg.static_penalty if k < cfg.n_static else cfg.dynamic_penalty
                    break
            self.acc += reward                                # :280
            done = self.goal_flag or self.hit                 # :286
        else:
            old = _dist(self.agent, self.goal)                # ballenv_pygame.py:652
            nx = self.agent[0] + action[0]
            ny = self.agent[1] + action[1]
            if nx < 0:
                nx = 0
            if nx > w:
                nx = w
            if ny < 0:
                ny = 0
            if ny > h:
                ny = h
            self.agent = (nx, ny)
            self.dist = _dist(self.agent, self.goal)          # :668
            self.goal_flag, self.hit, self.hit_index = False, False, -1
            for k, o in enumerate(self.obst):                 # :683-688
                if not (_dist(self.agent, o) - 0 > self.radius_sum):
                    self.hit, self.hit_index = True, k
                    break
            if self.hit:
                self.acc += -1
                reward, done = -1, True
            elif self.dist < self.goal_threshold:             # :690-697
                self.goal_flag = True
                self.acc += 1
                reward, done = 1, True
            else:                                             # :699-706
                reward = (old - self.dist) / self.total_distance
                self.acc += reward
                done = False
        self.tick += 1
        self.ep_len += 1
        return reward, done

    # ------------------------------------------------------------------ views
    def state(self):
        """[agent, goal, dist, obstacles...] like the reference's state list."""
        return [tuple(self.agent), tuple(self.goal), self.dist] + [tuple(o) for o in self.obst]

    def observe(self, window=None):
        return window_obs(self.agent, self.goal, self.obst, window or self.cfg.window, self.radius_sum)

    def observe_rows(self, window=None):
        return window_rows(self.agent, self.obst, window or self.cfg.window, self.radius_sum)


STAT_NAMES = ("episodes", "return_sum", "length_sum", "goals", "hits_static", "hits_dynamic",
              "timeouts", "steps")


class OracleVec:
    """N oracle envs with the vector wrapper's added semantics: the
    TimeLimit(1000) truncation (gym_ballenv/__init__.py:7), auto-reset of done
    envs (post-reset observation returned, terminal reward/done kept) and
    episode statistics."""

    def __init__(self, cfg: OracleConfig, source, n_envs: int, g0: int = 0):
        self.cfg = cfg
        self.envs = [OracleEnv(cfg, source, g0 + i) for i in range(n_envs)]
        self.stats = dict.fromkeys(STAT_NAMES, 0.0)

    def reset(self):
        for e in self.envs:
            e.reset()

    def reset_fixed(self, goal=(145, 120)):
        for e in self.envs:
            e.reset_fixed(goal)

    def step(self, actions):
        """actions: per env an index into AGENT_MOVES or a (dx, dy) pair.
        -> (rewards, dones, flags) ; flags bit0 goal, bit1 hit, bit2 truncated, bit3 hit is dynamic."""
        cfg = self.cfg
        rewards, dones, flags = [], [], []
        for e, a in zip(self.envs, actions):
            if not hasattr(a, "__len__"):
                a = AGENT_MOVES[int(a)]
            r, d = e.step(a)
            e.truncated = cfg.max_episode_steps > 0 and e.ep_len >= cfg.max_episode_steps
            done = d or e.truncated
            f = (1 if e.goal_flag else 0) | (2 if e.hit else 0) | (4 if e.truncated else 0)
            if e.hit and e.hit_index >= cfg.n_static:
                f |= 8
            s = self.stats
            s["steps"] += 1
            if done:
                s["episodes"] += 1
                s["return_sum"] += e.acc
                s["length_sum"] += e.ep_len
                s["goals"] += 1 if e.goal_flag else 0
                s["hits_static"] += 1 if (e.hit and not f & 8) else 0
                s["hits_dynamic"] += 1 if f & 8 else 0
                s["timeouts"] += 1 if (e.truncated and not d) else 0
                if cfg.auto_reset:
                    e.reset()
            rewards.append(r)
            dones.append(done)
            flags.append(f)
        return rewards, dones, flags

    def observe(self):
        return [e.observe() for e in self.envs]


# ---------------------------------------------------------------------------------------------------------------
# 20-float social-navigation features (featureExtractor.py:247-265), restated with math.* on Python floats.
def _angle_between(v1, v2):
    """featureExtractor.py:43-56: arccos(clip(dot(unit(v1), unit(v2)), -1, 1)); zero vectors stay zero."""
    n1, n2 = math.hypot(v1[0], v1[1]), math.hypot(v2[0], v2[1])
    u1 = (v1[0] / n1, v1[1] / n1) if n1 > 0 else v1
    u2 = (v2[0] / n2, v2[1] / n2) if n2 > 0 else v2
    d = u1[0] * u2[0] + u1[1] * u2[1]
    return math.acos(max(-1.0, min(1.0, d)))


def features20(agent, goal, obstacles, agent_rad=10, obstacle_rad=20, agent_vel=(0, 0), obstacle_vel=(0, 0)):
    """-> list of 20 floats: [goal-distance bin] + goal direction one-hot[4] + density[3] + orientation x speed
    histogram[9] (row-major, rows = orientation bin) + social forces[3]."""
    f = [0.0] * 20
    d = math.floor(math.hypot(agent[0] - goal[0], agent[1] - goal[1]) / 5)            # :132-144
    f[0] = 5.0 if d > 5 else float(d)
    xi, yi = goal[0] - agent[0], goal[1] - agent[1]                                    # :146-166
    ang = _angle_between((0, 1), (xi, yi))
    if ang < math.pi / 4:
        f[1] = 1.0
    elif ang > math.pi / 4 and ang < math.pi * 3 / 4:
        f[2 if xi > 0 else 4] = 1.0
    else:
        f[3] = 1.0
    rv = (obstacle_vel[0] - agent_vel[0], obstacle_vel[1] - agent_vel[1])
    relvel = math.hypot(rv[0], rv[1])
    speed_bin = 0 if relvel < 0.015 else (1 if relvel < 0.025 else 2)                  # :66-72
    for o in obstacles:
        v1 = (o[0] - agent[0], o[1] - agent[1])
        dist = math.hypot(v1[0], v1[1]) - agent_rad - obstacle_rad                     # :36-40
        if dist < 1000:                                                                # :103-108
            f[7] += 1
        if dist < 230:
            f[6] += 1
        if dist < 101:
            f[5] += 1
        psi = _angle_between(v1, rv)                                                   # :74
        obin = 0 if psi < math.pi / 4 else (1 if (psi > math.pi / 4 and psi < math.pi * 3 / 4) else 2)   # :76-83
        f[8 + obin * 3 + speed_bin] += 1                                               # :122-124
        fsoc = 1 * (1 * math.exp(-dist / 10)) * dist * (2 + 0.5 * (1 - 2) * (1 + math.cos(psi)))   # :186-189
        if fsoc > 1:                                                                   # :191-192
            f[17 + obin] += fsoc
    return f


def blocks29(agent, goal, obstacles):
    """-> list of 29 floats: the legacy observation of the REINFORCE / imitation scripts
    (examples/ball_env_reinforce.py:130-172, prep_state2 + block_to_arrpos): 4 goal-quadrant bits, then a 5 x 5 grid
    of obstacle COUNTS in 20-pixel blocks around the agent, the agent's own cell (index 4 + 12) starting at 1.

    x_block = sign(dx) * (dx - 10) // 20 with dx = agent_x - obstacle_x (and the same for y) unless dx or dy is 0, in
    which case both blocks are 0 (:153-157).  The formula is not symmetric: dx > 0 gives floor((dx - 10) / 20), dx < 0
    gives floor((|dx| + 10) / 20) - obstacles to the right of the agent never land in a negative block."""
    ref = [0.0] * 29
    ref[12 + 4] = 1.0                                                                  # :138
    ref[goal_quadrant(agent, goal)] = 1.0                                              # :139-150
    for o in obstacles:                                                                # :152 (state[3:], list order)
        xd, yd = agent[0] - o[0], agent[1] - o[1]
        xb = yb = 0
        if xd != 0 and yd != 0:                                                        # :156
            xb = math.floor((1 if xd > 0 else -1) * (xd - 10) / 20)                   # :157
            yb = math.floor((1 if yd > 0 else -1) * (yd - 10) / 20)                   # :158
        if abs(xb) < 3 and abs(yb) < 3:                                                # :160
            ref[4 + 12 + 5 * int(yb) + int(xb)] += 1.0                                 # :163-164, :169-172
    return ref
