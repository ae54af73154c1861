"""Environment configuration with the reference's field names.

The reference passes an argparse Namespace to BallEnv.customize_environment
(gym_ballenv/envs/ballenv_env.py:87-109); the fields and defaults below are those of
examples/ball_cnn_ac3.py:40-51.  ``EnvConfig.from_args`` accepts such a Namespace unchanged.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Sequence

from . import _lib as L


def _parse_goal(g):
    if isinstance(g, str):           # "x,y" strings, ballenv_env.py:99-103
        parts = g.strip().split(",")
        return int(parts[0]), int(parts[1])
    return float(g[0]), float(g[1])


@dataclass
class EnvConfig:
    static_obstacles: int = 13
    dynamic_obstacles: int = 5
    obstacle_speed: Sequence[float] = (1, 1, 1, 1, 1)
    obs_goal_position: Sequence = ('12,122', '123,93', '87,150', '430,440', '230,11')
    time_step_for_change: int = 50
    rd_th_obs: int = 60
    rd_th_agent: int = 80                      # stored, never used by the reference (:97)
    static_thresholds: Sequence[int] = (0, 0)   # stored, used only in commented-out code (:209-219)
    dynamic_thresholds: Sequence[int] = (10, 10)
    static_penalty: Sequence[float] = (1, 1)    # only index 1 is used (:138,223)
    dynamic_penalty: Sequence[float] = (4000, 8000)
    # pygame ruleset (createBoard ctor, ballenv_pygame.py:316)
    agent_radius: float = 10
    static_obstacle_radius: float = 10

    @classmethod
    def from_args(cls, args):
        """Build from the argparse Namespace the reference's scripts create (read_arguments())."""
        kw = {}
        for name in cls.__dataclass_fields__:
            if hasattr(args, name):
                kw[name] = getattr(args, name)
        cfg = cls(**kw)
        cfg.validate()
        return cfg

    @classmethod
    def pygame_default(cls, static_obstacles=0, agent_radius=10, static_obstacle_radius=10):
        return cls(static_obstacles=static_obstacles, dynamic_obstacles=0, obstacle_speed=(),
                   obs_goal_position=(), agent_radius=agent_radius, static_obstacle_radius=static_obstacle_radius)

    @classmethod
    def dense_moving(cls):
        """BASELINE.json config 3: 8 static + 24 moving obstacles, goals on a 6 x 4 lattice."""
        goals = ['%d,%d' % (x, y) for y in (100, 200, 300, 400) for x in (50, 130, 210, 290, 370, 450)]
        return cls(static_obstacles=8, dynamic_obstacles=24, obstacle_speed=[1] * 24, obs_goal_position=goals)

    def goals(self):
        return [_parse_goal(g) for g in self.obs_goal_position]

    def validate(self):
        # same checks as assert_arguments(), examples/ball_cnn_ac3.py:61-68
        if len(self.obstacle_speed) != self.dynamic_obstacles:
            raise ValueError("The length of the list of obstacle_speed does not match the no. of dynamic obstacles")
        if self.dynamic_obstacles and len(self.obs_goal_position) < self.dynamic_obstacles:
            raise ValueError("The length of the list of obstacle_goal_position does not match the no. of dynamic obstacles")
        for name in ("static_thresholds", "dynamic_thresholds", "static_penalty", "dynamic_penalty"):
            if len(getattr(self, name)) != 2:
                raise ValueError("The length of the list of %s is not equal to 2" % name)
        if self.dynamic_obstacles > L.MAX_DYNAMIC or len(self.obs_goal_position) > L.MAX_GOALS:
            raise ValueError("at most %d dynamic obstacles / %d goals" % (L.MAX_DYNAMIC, L.MAX_GOALS))

    def to_c(self, window, ruleset=L.RULESET_GYM, precision=L.F32, obs_format=L.OBS_F32,
             max_episode_steps=1000, auto_reset=True):
        self.validate()
        c = L.BallenvConfig()
        c.abi_version = L.ABI_VERSION
        c.ruleset = ruleset
        c.window = int(window)
        c.static_obstacles = int(self.static_obstacles)
        c.dynamic_obstacles = int(self.dynamic_obstacles)
        goals = self.goals()
        c.n_goals = len(goals)
        c.time_step_for_change = int(self.time_step_for_change)
        c.rd_th_obs = int(self.rd_th_obs)
        c.max_episode_steps = int(max_episode_steps)
        c.auto_reset = 1 if auto_reset else 0
        c.precision = precision
        c.obs_format = obs_format
        c.static_penalty = float(self.static_penalty[1])
        c.dynamic_penalty = float(self.dynamic_penalty[1])
        c.agent_radius = float(self.agent_radius)
        c.static_obstacle_radius = float(self.static_obstacle_radius)
        for j, s in enumerate(self.obstacle_speed):
            c.obstacle_speed[j] = float(s)        # the CLI passes strings when given (--obstacle_speed has no type=)
        for i, (x, y) in enumerate(goals):
            c.obs_goal_x[i], c.obs_goal_y[i] = float(x), float(y)
        return c
