"""CPU rollout worker for bench.py's ``cpu_baseline`` leg and ``--impl reference`` arm.

TEST / MEASUREMENT INFRASTRUCTURE ONLY (see oracle/ballenv_oracle.py).  The reference is pure Python
and cannot travel to the GPU box, so its CPU cost is timed through the oracle port: the same scalar
``math.sqrt``/``math.pow`` loops as ``BallEnv.step`` (gym_ballenv/envs/ballenv_env.py:232-289) and
``prep_state4`` (examples/ball_cnn_ac3.py:384-412), one Python process per host core - the
"multiprocessing loop on the box's own host cores" BASELINE.json asks for.  No torch import here,
so spawned workers start in milliseconds.
"""
from __future__ import annotations

import random
import time

from . import draws as D
from .ballenv_oracle import OracleConfig, OracleVec

_VEC = {}


def oracle_config(workload: dict) -> OracleConfig:
    """workload: {"window", "static_obstacles", "dynamic_obstacles", "goals", "speeds"} (plain types)."""
    return OracleConfig(window=workload["window"], n_static=workload["static_obstacles"],
                        n_dynamic=workload["dynamic_obstacles"], speeds=list(workload["speeds"]),
                        goals=[tuple(g) for g in workload["goals"]], change_step=workload.get("change_step", 50),
                        rd_th_obs=workload.get("rd_th_obs", 60), max_episode_steps=1000, auto_reset=True)


def _vec(workload, worker, n_envs):
    key = (worker, n_envs, workload["window"], workload["static_obstacles"], workload["dynamic_obstacles"])
    if key not in _VEC:
        v = OracleVec(oracle_config(workload), D.PhiloxDraws(0), n_envs, g0=worker * n_envs)
        v.reset()
        _VEC[key] = (v, random.Random(1 + worker))
    return _VEC[key]


def run_steps(args):
    """Advance ``n_envs`` oracle envs by ``n_steps`` step+observe each -> (env_steps, seconds)."""
    workload, worker, n_envs, n_steps = args
    v, rng = _vec(workload, worker, n_envs)
    t0 = time.perf_counter()
    for _ in range(n_steps):
        v.step([rng.randrange(9) for _ in range(n_envs)])
        v.observe()
    return n_envs * n_steps, time.perf_counter() - t0


def run_for(args):
    """Step+observe until ``seconds`` have elapsed -> (env_steps, seconds)."""
    workload, worker, n_envs, seconds = args
    v, rng = _vec(workload, worker, n_envs)
    t0 = time.perf_counter()
    n = 0
    while time.perf_counter() - t0 < seconds:
        v.step([rng.randrange(9) for _ in range(n_envs)])
        v.observe()
        n += n_envs
    return n, time.perf_counter() - t0
