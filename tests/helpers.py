"""Shared test helpers: golden-fixture loading and config translation."""
import json
import os

import numpy as np

from oracle import draws as D
from oracle.ballenv_oracle import OracleConfig, RULESET_GYM

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = json.loads(str(z["meta"]))
    return z, meta


def parse_goals(cfg):
    return [tuple(int(v) for v in s.split(",")) for s in cfg["obs_goal_position"]]


def oracle_config(cfg, window, max_episode_steps=1000, auto_reset=True):
    return OracleConfig(
        ruleset=RULESET_GYM, window=window, n_static=cfg["static_obstacles"], n_dynamic=cfg["dynamic_obstacles"],
        speeds=list(cfg["obstacle_speed"]), goals=parse_goals(cfg), change_step=cfg["time_step_for_change"],
        rd_th_obs=cfg["rd_th_obs"], static_penalty=cfg["static_penalty"][1], dynamic_penalty=cfg["dynamic_penalty"][1],
        max_episode_steps=max_episode_steps, auto_reset=auto_reset)


def tapes_from_golden(z, meta):
    """Rebuild uint32 tape words from the recorded (value, n) pairs of an mt-mode rollout."""
    cfg = meta["cfg"]
    ks, kd, A = cfg["static_obstacles"], cfg["dynamic_obstacles"], meta["tape_attempts"]
    val, n = z["tape_step_val"].astype(np.uint64), z["tape_step_n"].astype(np.uint64)
    step = np.zeros(val.shape, np.uint32)
    nz = n > 0
    step[nz] = ((val[nz] << np.uint64(32)) + n[nz] - np.uint64(1)) // n[nz]     # ceil(v * 2^32 / n)
    rv = z["tape_reset_val"].astype(np.uint64)
    width = D.reset_tape_width(ks, kd, A)
    ns = np.zeros(width, np.uint64)
    ns[0:4] = (500, 20, 500, 10)
    ns[4::2] = 500
    ns[5::2] = 460
    reset = (((rv << np.uint64(32)) + ns - np.uint64(1)) // ns).astype(np.uint32)
    return step, reset


def canonical_goal_index(cfg):
    """index -> first index of a goal with the same position.  The reference holds an obstacle's goal as a position
    (ballenv_env.py:351-353), so with repeated goals an index is only defined up to that mapping."""
    goals = parse_goals(cfg)
    return np.array([goals.index(g) for g in goals])
