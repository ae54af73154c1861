"""gym_ballenv_b200 - B200-native batched step()/reset()/window-observe of ranok92/gym-ballenv.

Public surface
  BallVecEnv              N environments per launch, CUDA tensors in/out (new; vec_env.py)
  BallEnv, make           the registered gym env's API on the same kernels (env.py)
  createBoard             the stand-alone pygame env's API (ballenv_pygame.py) incl. sensor_readings (env.py)
  EnvConfig               the reference's argparse fields as a dataclass (config.py)
  make_prep_state         prep_state2 / prep_state4 of examples/ball_cnn_ac3.py, GPU-evaluated
  make_sharded_env, shard_bounds, allreduce_stats   one process per GPU, envs partitioned by global id (distributed.py)
Importing this package loads libballenv_b200.so (built by ``python -m gym_ballenv_b200.build``);
there is no CPU fallback.
"""
from ._lib import (LIB, BallenvError, FLAG_GOAL, FLAG_HIT, FLAG_HIT_DYNAMIC, FLAG_TRUNCATED, LIB_PATH,
                   STAT_NAMES)
from .config import EnvConfig
from .env import BallEnv, TimeLimit, createBoard, make, make_prep_state
from .vec_env import MOVE_LIST, BallVecEnv
from .distributed import allreduce_stats, make_sharded_env, shard_bounds
from . import legacy, pathlogs

__all__ = ["legacy", "pathlogs", "BallVecEnv", "BallEnv", "createBoard", "TimeLimit", "make", "make_prep_state", "EnvConfig", "MOVE_LIST",
           "allreduce_stats", "make_sharded_env", "shard_bounds", "BallenvError", "FLAG_GOAL", "FLAG_HIT", "FLAG_TRUNCATED", "FLAG_HIT_DYNAMIC", "STAT_NAMES", "LIB_PATH"]


def _register_with_gym():
    """If a real ``gym`` is importable, register the id the reference registers - ``gymball-v0``
    (gym_ballenv/__init__.py:4-11: entry point, max_episode_steps=1000, reward_threshold=100.0) - pointing at this
    implementation, so that a script calling ``gym.make('gymball-v0')`` is served unchanged.  Skipped when the reference
    package has registered the id already, or when BALLENV_NO_GYM_REGISTER=1.  (Without gym, ``gym_ballenv_b200.make``
    takes the same id.)"""
    import os
    if os.environ.get("BALLENV_NO_GYM_REGISTER", "0") == "1":
        return False
    try:
        import gym
        from gym.envs.registration import register
    except ImportError:
        return False
    try:
        register(id='gymball-v0', entry_point='gym_ballenv_b200.env:BallEnv', max_episode_steps=1000,
                 reward_threshold=100.0, nondeterministic=False)
    except gym.error.Error:      # the id is taken (the reference package was imported first): leave it alone
        return False
    return True


_register_with_gym()
