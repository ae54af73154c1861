"""ctypes wrapper of oracle/ballenv_oracle.c (libballenv_oracle.so, built by oracle/Makefile).

TEST INFRASTRUCTURE ONLY - see the header of ballenv_oracle.c.  Same interface shape as
``oracle.ballenv_oracle.OracleVec`` with numpy arrays instead of lists.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libballenv_oracle.so")
STAT_NAMES = ("episodes", "return_sum", "length_sum", "goals", "hits_static", "hits_dynamic", "timeouts", "steps")


class OrcConfig(C.Structure):
    _fields_ = [("window", C.c_int32), ("n_static", C.c_int32), ("n_dynamic", C.c_int32), ("n_goals", C.c_int32),
                ("change_step", C.c_int32), ("rd_th_obs", C.c_int32), ("max_episode_steps", C.c_int32),
                ("auto_reset", C.c_int32), ("static_penalty", C.c_double), ("dynamic_penalty", C.c_double),
                ("speeds", C.c_double * 64), ("goal_x", C.c_double * 64), ("goal_y", C.c_double * 64)]


def build():
    src = os.path.join(HERE, "ballenv_oracle.c")
    if not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "-s"], check=True)
    return LIB_PATH


_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        l = C.CDLL(build())
        l.orc_create.restype = C.c_void_p
        l.orc_create.argtypes = [C.POINTER(OrcConfig), C.c_int64, C.c_uint64, C.c_int64]
        l.orc_destroy.argtypes = [C.c_void_p]
        l.orc_reset.argtypes = [C.c_void_p]
        l.orc_step.argtypes = [C.c_void_p] + [C.c_void_p] * 4
        l.orc_observe.argtypes = [C.c_void_p, C.c_void_p]
        l.orc_stats.argtypes = [C.c_void_p, C.c_void_p]
        l.orc_get_state.argtypes = [C.c_void_p] + [C.c_void_p] * 4
        _LIB = l
    return _LIB


class COracleVec:
    """N environments of the gym ruleset with Philox draws (production addressing), auto-reset and TimeLimit."""

    def __init__(self, cfg, seed: int, n_envs: int, g0: int = 0):
        """cfg: oracle.ballenv_oracle.OracleConfig (gym ruleset)."""
        assert cfg.ruleset == 0, "the C oracle restates the gym ruleset"
        c = OrcConfig()
        c.window, c.n_static, c.n_dynamic = cfg.window, cfg.n_static, cfg.n_dynamic
        c.n_goals = len(cfg.goals) if cfg.n_dynamic else 0
        c.change_step, c.rd_th_obs = cfg.change_step, cfg.rd_th_obs
        c.max_episode_steps, c.auto_reset = cfg.max_episode_steps, 1 if cfg.auto_reset else 0
        c.static_penalty, c.dynamic_penalty = float(cfg.static_penalty), float(cfg.dynamic_penalty)
        for j in range(cfg.n_dynamic):
            c.speeds[j] = float(cfg.speeds[j])
        for i in range(c.n_goals):
            c.goal_x[i], c.goal_y[i] = float(cfg.goals[i][0]), float(cfg.goals[i][1])
        self.cfg, self.n = cfg, int(n_envs)
        self.row = 4 + cfg.window * cfg.window
        self._l = lib()
        self._h = self._l.orc_create(C.byref(c), self.n, C.c_uint64(seed & (2 ** 64 - 1)), int(g0))
        assert self._h

    def __del__(self):
        if getattr(self, "_h", None):
            self._l.orc_destroy(self._h)
            self._h = None

    def reset(self):
        self._l.orc_reset(self._h)

    def step(self, actions):
        """actions: int array [n] of indices into the agent move table -> (reward f64, done bool, flags u8)."""
        a = np.ascontiguousarray(actions, dtype=np.int64)
        assert a.shape == (self.n,) and a.min() >= 0 and a.max() <= 8
        r = np.empty(self.n, np.float64)
        d = np.empty(self.n, np.uint8)
        f = np.empty(self.n, np.uint8)
        self._l.orc_step(self._h, a.ctypes.data, r.ctypes.data, d.ctypes.data, f.ctypes.data)
        return r, d.astype(bool), f

    def observe(self):
        o = np.empty((self.n, self.row), np.float32)
        self._l.orc_observe(self._h, o.ctypes.data)
        return o

    @property
    def stats(self):
        out = np.zeros(8, np.float64)
        self._l.orc_stats(self._h, out.ctypes.data)
        return dict(zip(STAT_NAMES, out.tolist()))

    def state(self):
        K, Kd = self.cfg.n_static + self.cfg.n_dynamic, self.cfg.n_dynamic
        s = np.empty((self.n, 7), np.float64)
        i = np.empty((self.n, 3), np.int64)
        o = np.empty((self.n, max(K, 1), 2), np.float64)
        d = np.empty((self.n, max(Kd, 1), 2), np.int32)
        self._l.orc_get_state(self._h, s.ctypes.data, i.ctypes.data, o.ctypes.data if K else None,
                              d.ctypes.data if Kd else None)
        return dict(agent=s[:, 0:2], goal=s[:, 2:4], dist=s[:, 4], total=s[:, 5], acc=s[:, 6], ep_len=i[:, 0],
                    episode=i[:, 1], tick=i[:, 2], obstacles=o[:, :K], dyn_goal=d[:, :Kd, 0], dyn_counter=d[:, :Kd, 1])
