"""Single-environment facade with the reference's gym API, running on the CUDA kernels.

Drop-in for ``gym.make('gymball-v0')`` (gym_ballenv/__init__.py:4-11): ``make()`` returns a
TimeLimit(1000)-wrapped ``BallEnv`` whose ``reset()`` / ``step(action)`` / ``customize_environment(args)``
/ ``unwrapped`` / attribute surface is the one the reference's training loops use
(examples/ball_cnn_ac3.py:468-489,541-613; examples/ball_dqn.py:43-45; examples/ball_env_reinforce.py).
Internally it is a one-environment ``BallVecEnv`` in fp64 parity mode, so the returned state list holds
the same numbers the reference's Python arithmetic produces.  There is no CPU fallback.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import numpy as np
import torch

from .config import EnvConfig
from .vec_env import BallVecEnv

_screen_width = 500
_screen_height = 500


def _num(v):
    v = float(v)
    return int(v) if v.is_integer() else v


class BallEnv(object):
    """gym_ballenv/envs/ballenv_env.py:37-389 re-hosted on the GPU (rendering excluded)."""

    metadata = {'render.modes': ['human', 'rgb_array'], 'video.frames_per_second': 100}

    def __init__(self, device="cuda", seed=0, window=5):
        # attributes the reference's callers read (ballenv_env.py:47-69)
        self.radius_rand_person = 20
        self.radius_ctrl_person = 5
        self.speed_rand_person = 1
        self.speedx_ctrl_person = 1
        self.speedy_ctrl_person = 1
        self.threshold_goal = 10
        self.timepenalty = 0
        self.total_reward_accumulated = 0
        self.reward_threshold = -20000
        self.old_dist = None
        self.total_distance = None
        self.framecount = 0
        self.viewer = None
        self.state = None
        self.action_space = SimpleNamespace(n=4)           # spaces.Discrete(4), vestigial (:57)
        self.observation_space = SimpleNamespace(low=np.array([0, 0]), high=np.array([_screen_width, _screen_height]))
        self.goal_x = self.goal_y = None
        self.no_of_static_obstacles = 0
        self.no_of_dynamic_obstacles = 0
        self.total_obstacles = 0
        self._device = device
        self._seed = seed
        self._window = window
        self._cfg = EnvConfig(static_obstacles=0, dynamic_obstacles=0, obstacle_speed=(), obs_goal_position=())
        self._vec = None
        self._obs = None
        self._scratch = {}

    @property
    def unwrapped(self):
        return self

    # ------------------------------------------------------------------ configuration
    def customize_environment(self, args):
        """ballenv_env.py:87-109: ``args`` is the argparse Namespace of the example scripts (or an EnvConfig)."""
        self._cfg = args if isinstance(args, EnvConfig) else EnvConfig.from_args(args)
        c = self._cfg
        self.no_of_static_obstacles = c.static_obstacles
        self.no_of_dynamic_obstacles = c.dynamic_obstacles
        self.obstacle_speed_list = list(c.obstacle_speed)
        self.goal_change_step = c.time_step_for_change
        self.obs_uncertainity_threshold = c.rd_th_obs
        self.agent_uncertainty_threshold = c.rd_th_agent
        self.obstacle_goal_list = [tuple(g) for g in c.goals()]
        self.total_obstacles = c.static_obstacles + c.dynamic_obstacles
        self.st_obs_prox_thresh, self.dy_obs_prox_thresh = c.static_thresholds, c.dynamic_thresholds
        self.st_obs_prox_penalty, self.dy_obs_prox_penalty = c.static_penalty, c.dynamic_penalty
        if self._vec is not None:
            self._vec.close()
        self._vec = None

    def set_window(self, window):
        """WINDOW of examples/ball_cnn_ac3.py:493 for the fused observation returned by window_observation()."""
        if window != self._window:
            self._window = window
            if self._vec is not None:
                self._vec.close()
            self._vec = None

    def seed(self, seed=None):
        """ballenv_env.py:174-177.  Here the seed actually keys the env's Philox streams (takes effect at the next reset
        of a freshly configured env)."""
        if seed is not None:
            self._seed = int(seed)
            if self._vec is not None:
                self._vec.close()
            self._vec = None
        return [self._seed]

    def _ensure(self):
        if self._vec is None:
            self._vec = BallVecEnv(1, window=self._window, config=self._cfg, ruleset="gym", device=self._device,
                                   seed=self._seed, parity=True, auto_reset=False, max_episode_steps=0)
        return self._vec

    # ------------------------------------------------------------------ reference API
    def reset(self):
        """ballenv_env.py:113-167 -> object ndarray [agent, goal, dist, obstacles...]."""
        v = self._ensure()
        self._obs = v.reset()
        self.total_reward_accumulated = 0
        return self._pull_state()

    def step(self, action):
        """ballenv_env.py:232-289 -> (state, reward, done, {})."""
        self.framecount += 1
        v = self._ensure()
        if self.state is None:
            raise RuntimeError("call reset() before step()")
        a = torch.tensor([[float(action[0]), float(action[1])]], dtype=torch.float64, device=v.device)
        self._obs, reward, done, _ = v.step(a)
        reward = float(reward.item())
        done = bool(done.item())
        state = self._pull_state()
        return state, reward, done, {}

    def _pull_state(self):
        v = self._vec
        sv = v.state_views      # fp64 (parity mode): the one environment's slice, gathered on the device, one small copy
        ks, kd = self._cfg.static_obstacles, self._cfg.dynamic_obstacles
        flat = torch.cat([sv[k][0:1] for k in ("agent_x", "agent_y", "goal_x", "goal_y", "dist", "total_distance",
                                               "acc_reward")] +
                         [sv[k][:, 0] for k in ("static_x", "static_y", "dynamic_x", "dynamic_y")]).cpu().tolist()
        ax, ay, gx, gy, dist = flat[0], flat[1], flat[2], flat[3], float(flat[4])
        self.total_distance = float(flat[5])
        self.total_reward_accumulated = float(flat[6])
        state = [(_num(ax), _num(ay)), (_num(gx), _num(gy)), dist]
        o = 7
        if ks:
            sx, sy = flat[o:o + ks], flat[o + ks:o + 2 * ks]
            state += [(_num(x), _num(y)) for x, y in zip(sx, sy)]
        o += 2 * ks
        if kd:
            dx, dy = flat[o:o + kd], flat[o + kd:o + 2 * kd]
            state += [(float(x), float(y)) for x, y in zip(dx, dy)]
        self.goal_x, self.goal_y = state[1]
        self.old_dist = dist
        self.state = state
        out = np.empty(len(state), dtype=object)
        for i, s in enumerate(state):
            out[i] = s
        return out

    def get_accumulated_reward(self):
        return self.total_reward_accumulated

    def calculate_distance(self, tup1, tup2):
        return math.sqrt(math.pow(tup1[0] - tup2[0], 2) + math.pow(tup1[1] - tup2[1], 2))

    def check_overlap(self, tup1, tup2):
        """ballenv_env.py:185-191 (host-side helper kept for callers; the kernels do this test on the device)."""
        return not (self.calculate_distance(tup1, tup2) > (self.radius_rand_person + self.radius_ctrl_person))

    def check_overlap_rect(self, tup1, tup2, rad):
        return (abs(tup1[0] - tup2[0]) < (rad + self.radius_ctrl_person)
                and abs(tup1[1] - tup2[1]) < (rad / 2 + self.radius_ctrl_person))

    def render(self, mode='rgb_array', close=True):
        """Rendering stays off the hot path (pyglet viewer of ballenv_env.py:357-386 is not rebuilt)."""
        return None

    def close(self):
        if self._vec is not None:
            self._vec.close()
            self._vec = None
        for v in self._scratch.values():
            v.close()
        self._scratch = {}

    # ------------------------------------------------------------------ fused observation
    def window_observation(self):
        """prep_state4(state, WINDOW) of the current state, computed by the step/reset launch itself:
        float32 [1, 4 + W*W] on the device (examples/ball_cnn_ac3.py:412)."""
        return self._obs

    def observe_state(self, state, window):
        """prep_state4 for an arbitrary reference-style state list (any number of obstacles), on the GPU."""
        k = len(state) - 3
        key = (window, k)
        v = self._scratch.get(key)
        if v is None:
            cfg = EnvConfig(static_obstacles=k, dynamic_obstacles=0, obstacle_speed=(), obs_goal_position=())
            v = BallVecEnv(1, window=window, config=cfg, device=self._device, parity=True, auto_reset=False,
                           max_episode_steps=0)
            self._scratch[key] = v
        fields = dict(agent_x=[state[0][0]], agent_y=[state[0][1]], goal_x=[state[1][0]], goal_y=[state[1][1]])
        if k:
            fields["static_x"] = [[float(p[0])] for p in state[3:]]
            fields["static_y"] = [[float(p[1])] for p in state[3:]]
        v.set_state(**fields)
        return v.observe()


    def extract_patch(self, state=None, width=100, interp="bicubic"):
        """``extract_patch(state, width)`` of the pixel-policy scripts (examples/ball_cnn_reinforce.py:120-163): the
        viewer's frame around the agent, resized to 40 x 40 - float32 [1, 3, 40, 40] in [0, 1] on the device, rendered
        by the GPU (BallVecEnv.rgb_patches) instead of pyglet + PIL.  ``state``: a reference-style state list, or None
        for the current state."""
        if state is None or (self.state is not None and len(state) == len(self.state)
                             and all(a == b for a, b in zip(state, self.state))):
            self._ensure()
            return self._vec.rgb_patches(width=width, interp=interp)
        k = len(state) - 3
        ks = self._cfg.static_obstacles if k == self._cfg.static_obstacles + self._cfg.dynamic_obstacles else k
        key = ("patch", ks, k - ks)
        v = self._scratch.get(key)
        if v is None:
            cfg = EnvConfig(static_obstacles=ks, dynamic_obstacles=k - ks,
                            obstacle_speed=tuple(self._cfg.obstacle_speed)[:k - ks] if k > ks else (),
                            obs_goal_position=self._cfg.obs_goal_position if k > ks else ())
            v = BallVecEnv(1, window=5, config=cfg, device=self._device, parity=True, auto_reset=False, max_episode_steps=0)
            self._scratch[key] = v
        fields = dict(agent_x=[state[0][0]], agent_y=[state[0][1]], goal_x=[state[1][0]], goal_y=[state[1][1]])
        if ks:
            fields["static_x"] = [[float(p[0])] for p in state[3:3 + ks]]
            fields["static_y"] = [[float(p[1])] for p in state[3:3 + ks]]
        if k > ks:
            fields["dynamic_x"] = [[float(p[0])] for p in state[3 + ks:]]
            fields["dynamic_y"] = [[float(p[1])] for p in state[3 + ks:]]
        v.set_state(**fields)
        return v.rgb_patches(width=width, interp=interp)


class TimeLimit(object):
    """The wrapper gym.make() puts around the env for ``timestep_limit = 1000`` (gym_ballenv/__init__.py:7;
    gym 0.10.9 ``TimeLimit``: done once elapsed steps >= max_episode_steps)."""

    def __init__(self, env, max_episode_steps=1000):
        self.env = env
        self._max_episode_steps = max_episode_steps
        self._elapsed_steps = 0

    @property
    def unwrapped(self):
        return self.env.unwrapped

    def __getattr__(self, name):
        return getattr(self.env, name)

    def reset(self):
        self._elapsed_steps = 0
        return self.env.reset()

    def step(self, action):
        state, reward, done, info = self.env.step(action)
        self._elapsed_steps += 1
        if self._elapsed_steps >= self._max_episode_steps:
            done = True
        return state, reward, done, info


def make(env_id='gymball-v0', **kwargs):
    """gym.make('gymball-v0') without gym."""
    if env_id != 'gymball-v0':
        raise ValueError("unknown environment id %r (only 'gymball-v0' is registered)" % env_id)
    return TimeLimit(BallEnv(**kwargs), max_episode_steps=1000)


def make_prep_state(env):
    """(prep_state2, prep_state4) with the signatures of examples/ball_cnn_ac3.py:330,384, evaluated on the GPU."""
    base = env.unwrapped

    def prep_state4(state, window):
        if state is not None and base.state is not None and window == base._window and base._obs is not None \
                and len(state) == len(base.state) and all(a == b for a, b in zip(state, base.state)):
            return base.window_observation()
        return base.observe_state(list(state), window)

    def prep_state2(state):
        return prep_state4(state, base._window)[:, 0:4]

    return prep_state2, prep_state4


class createBoard(object):
    """The stand-alone pygame environment of the reference (ballenv_pygame.py:314-706) re-hosted on the GPU:
    100 x 100 float world, static obstacles, raw (dx, dy) actions, hit => -1 (before the goal test), goal (< 15) => +1,
    otherwise the progress reward.  Same constructor keywords as the reference (:316); rendering, mouse / keyboard
    input and the dynamic-obstacle branch (undefined names in the reference, :502-506) are not rebuilt.
    ``sensor_readings`` is the 20-float feature vector of featureExtractor.py:247-265 as a float32 [1, 20] CUDA tensor."""

    def __init__(self, height=100, display=False, width=100, agent_radius=10, static_obstacles=0, dynamic_obstacles=0,
                 static_obstacle_radius=10, dynamic_obstacle_radius=0, obstacle_speed_list=(), device="cuda", seed=0,
                 window=5):
        if display:
            raise NotImplementedError("rendering stays off the hot path (pygame display is not rebuilt)")
        if dynamic_obstacles:
            raise ValueError("createBoard's dynamic obstacles are broken in the reference (ballenv_pygame.py:502-506)")
        if (height, width) != (100, 100):
            raise ValueError("the reference clamps to the module constants 100 x 100 (ballenv_pygame.py:8-9, 656-663)")
        self.height, self.width, self.display = height, width, display
        self.agent_radius = agent_radius
        self.no_static_obstacles, self.no_dynamic_obstacles = static_obstacles, 0
        self.total_obs = static_obstacles
        self.rad_static_obstacles, self.rad_dynamic_obstacles = static_obstacle_radius, dynamic_obstacle_radius
        self.goal_threshold = 15
        self.agent_x_vel = self.agent_y_vel = 0
        self.actionArray = [np.asarray([0, -1]), np.asarray([1, 0]), np.asarray([0, 1]), np.asarray([-1, 0])]   # :352
        self.state = self.sensor_readings = self.reward = None
        self.total_reward_accumulated = self.total_distance = self.old_dist = None
        self.agent_x = self.agent_y = self.goal_x = self.goal_y = None
        cfg = EnvConfig.pygame_default(static_obstacles=static_obstacles, agent_radius=agent_radius,
                                       static_obstacle_radius=static_obstacle_radius)
        self._vec = BallVecEnv(1, window=window, config=cfg, ruleset="pygame", device=device, seed=seed, parity=True,
                               auto_reset=False, max_episode_steps=0)

    def calculate_distance(self, tup1, tup2):
        return math.sqrt(math.pow(tup1[0] - tup2[0], 2) + math.pow(tup1[1] - tup2[1], 2))

    def check_overlap(self, tup1, tup2, thresh=0):
        return not (self.calculate_distance(tup1, tup2) - thresh > (self.rad_static_obstacles + self.agent_radius))

    def _pull(self):
        v = self._vec
        sv = v.state_views      # fp64: the one environment's slice, gathered on the device, one small copy
        ks = self.no_static_obstacles
        flat = torch.cat([sv[k][0:1] for k in ("agent_x", "agent_y", "goal_x", "goal_y", "dist", "total_distance",
                                               "acc_reward")] + [sv["static_x"][:, 0], sv["static_y"][:, 0]]).cpu().tolist()
        agent, goal = (flat[0], flat[1]), (flat[2], flat[3])
        self.agent_x, self.agent_y = agent
        self.goal_x, self.goal_y = goal
        self.total_distance = flat[5]
        self.total_reward_accumulated = flat[6]
        state = [agent, goal, flat[4]]
        # obstacle tuples carry the Obstacle.rad default 20 (ballenv_pygame.py:33-36, 496)
        state += [(int(x), int(y), 20) for x, y in zip(flat[7:7 + ks], flat[7 + ks:7 + 2 * ks])]
        self.state = state
        self.sensor_readings = v.sensor_readings()
        out = np.empty(len(state), dtype=object)
        for i, s in enumerate(state):
            out[i] = s
        return out

    def reset(self):
        """ballenv_pygame.py:460-513."""
        self._vec.reset()
        out = self._pull()
        self.old_dist = self.state[2]
        return out

    def resetFixedstate(self):
        """ballenv_pygame.py:589-624: goal at (145, 120), the obstacles of the last reset() kept, the agent redrawn until
        it is clear of them."""
        if self.state is None:
            raise RuntimeError("call reset() before resetFixedstate() (it keeps the obstacles of the last reset)")
        self._vec.reset_fixed((145, 120))
        out = self._pull()
        self.old_dist = self.state[2]
        return out

    def step(self, action):
        """ballenv_pygame.py:650-675 -> (state, reward, done, {})."""
        if self.state is None:
            raise RuntimeError("call reset() before step()")
        self.old_dist = self.calculate_distance(self.state[0], self.state[1])
        a = torch.tensor([[float(action[0]), float(action[1])]], dtype=torch.float64, device=self._vec.device)
        _, reward, done, _ = self._vec.step(a)
        reward, done = float(reward.item()), bool(done.item())
        return self._pull(), reward, done, {}

    def close(self):
        self._vec.close()
