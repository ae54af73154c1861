"""Loader that runs the UNMODIFIED reference code as the parity oracle's pin.

TEST INFRASTRUCTURE ONLY (see oracle/draws.py header).  Works only where
``/root/reference`` exists (the build container); the GPU box uses the golden
fixtures this produces (``oracle/gen_golden.py`` -> ``tests/golden/*.npz``).

Recipe (no reference file is edited or copied):
 1. a stand-in ``gym`` package is put in ``sys.modules`` providing only what
    ``gym_ballenv/envs/ballenv_env.py:1-7`` and ``gym_ballenv/__init__.py:1``
    touch;
 2. ``gym_ballenv.envs.ballenv_env`` is imported from ``/root/reference``;
 3. the module global ``np`` is replaced by a proxy that (a) lets
    ``np.array(ragged_state)`` fall back to ``dtype=object`` (NumPy >= 1.24
    raises; the reference relied on NumPy 1.15's implicit object arrays,
    ``ballenv_env.py:167,289``) and (b) routes ``np.random.randint`` to a
    draw *router* that reconstructs the address of every draw;
 4. ``prep_state2`` / ``prep_state4`` are lifted out of
    ``examples/ball_cnn_ac3.py:330-352,384-412`` by AST (the script itself
    cannot be imported: top-level argparse / gym.make / matplotlib);
 5. ``static_obstacle_list`` / ``dynamic_obstacle_list`` are cleared before
    every ``reset()`` (the reference never clears them, ``ballenv_env.py:114,
    146,162``; stale entries only burn RNG draws and time).
"""
from __future__ import annotations

import ast
import math
import os
import sys
import types
from argparse import Namespace

import numpy as _np

from . import draws as D

REFERENCE_ROOT = os.environ.get("BALLENV_REFERENCE", "/root/reference")


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "gym_ballenv", "envs", "ballenv_env.py"))


# --------------------------------------------------------------------------- gym stand-in
def _install_gym_standin():
    if "gym" in sys.modules and not getattr(sys.modules["gym"], "_ballenv_standin", False):
        return  # a real gym is importable: use it
    gym = types.ModuleType("gym")
    gym._ballenv_standin = True

    class Env(object):
        metadata = {}

        @property
        def unwrapped(self):
            return self

    class _Space(object):
        def __init__(self, *a, **k):
            self.args = a
            self.n = a[0] if a and isinstance(a[0], int) else None

    gym.Env = Env
    gym.error = types.ModuleType("gym.error")
    gym.spaces = types.ModuleType("gym.spaces")
    gym.spaces.Discrete = _Space
    gym.spaces.Box = _Space
    gym.utils = types.ModuleType("gym.utils")
    gym.utils.seeding = types.ModuleType("gym.utils.seeding")
    gym.utils.seeding.np_random = lambda seed=None: (_np.random.RandomState(seed), seed)
    gym.envs = types.ModuleType("gym.envs")
    gym.envs.registration = types.ModuleType("gym.envs.registration")
    gym.registry = {}
    gym.envs.registration.register = lambda **kw: gym.registry.__setitem__(kw["id"], kw)
    for name in ("error", "spaces", "utils", "utils.seeding", "envs", "envs.registration"):
        obj = gym
        for part in name.split("."):
            obj = getattr(obj, part)
        sys.modules["gym." + name] = obj
    sys.modules["gym"] = gym


# --------------------------------------------------------------------------- numpy proxy
class _RandomProxy(object):
    def __init__(self, router):
        self._router = router

    def randint(self, low, high=None, *a, **k):
        if high is None:
            low, high = 0, low
        return self._router.randint(int(low), int(high))

    def ranf(self, *a, **k):
        return self._router.ranf()

    def __getattr__(self, name):
        return getattr(_np.random, name)


class _NumpyProxy(object):
    def __init__(self, router):
        self.random = _RandomProxy(router)

    def array(self, obj, *a, **k):
        try:
            return _np.array(obj, *a, **k)
        except ValueError:          # ragged state list -> object array (NumPy 1.15 behaviour)
            out = _np.empty(len(obj), dtype=object)
            for i, v in enumerate(obj):
                out[i] = v
            return out

    asarray = array

    def __getattr__(self, name):
        return getattr(_np, name)


# --------------------------------------------------------------------------- draw routers
class MTRouter(object):
    """Forwards to the real global ``np.random`` (MT19937) and records
    (context, low, high, value) for every draw."""

    def __init__(self):
        self.log = []
        self.context = "ctor"

    def randint(self, low, high):
        v = int(_np.random.randint(low, high))
        self.log.append((self.context, low, high, v))
        return v

    def ranf(self):
        v = float(_np.random.ranf())
        self.log.append((self.context, 0.0, 1.0, v))
        return v


class AddressedRouter(object):
    """Answers the gym-ruleset reference's draws from an addressed draw source
    (oracle.draws.PhiloxDraws / TapeDraws) by reconstructing, call by call,
    which draw the reference is asking for.

    reset (ballenv_env.py:113-164): goal_x, goal_y, agent_x, agent_y, then per
    obstacle an (x, y) pair; the static obstacle index is the number accepted
    so far (len(env.obstacle_list)), the attempt the number of pairs drawn
    while that index did not change.
    step (ballenv_env.py:323-353): per dynamic obstacle, draw A then, only if
    A was a ``randint(100)`` that came out >= rd_th_obs, draw B.
    """

    def __init__(self, source, g):
        self.src = source
        self.g = g
        self.env = None
        self.context = "ctor"
        self.episode = -1
        self.tick = 0
        self.trace = []      # (context, low, high, value)

    # harness hooks
    def begin_reset(self):
        self.context = "reset"
        self.episode += 1
        self._n = 0
        self._pair = None
        self._last_len = -1
        self._attempt = 0

    def begin_step(self):
        self.context = "step"
        self._j = 0
        self._second = None

    def end_step(self):
        self.tick += 1

    def randint(self, low, high):
        n = high - low
        if self.context == "ctor":
            w = 0
        elif self.context == "reset":
            w = self._reset_word()
        else:
            w = self._step_word(n)
        v = low + D.mulhi(w, n)
        self.trace.append((self.context, low, high, v))
        return v

    def _reset_word(self):
        k = self._n
        self._n += 1
        if k < 4:
            return self.src.reset_words(self.g, self.episode, D.RK_HEAD, count=4)[k]
        env = self.env
        if self._pair is None:                      # x of a new (x, y) pair
            cur = len(env.obstacle_list)
            if cur != self._last_len:
                self._last_len, self._attempt = cur, 0
            else:
                self._attempt += 1
            if cur < env.no_of_static_obstacles:
                self._pair = self.src.reset_words(self.g, self.episode, D.RK_STATIC, cur, self._attempt)
            else:
                self._pair = self.src.reset_words(self.g, self.episode, D.RK_DYNAMIC,
                                                  cur - env.no_of_static_obstacles)
            return self._pair[0]
        w = self._pair[1]
        self._pair = None
        return w

    def _step_word(self, n):
        if self._second is not None:                # draw B of the same obstacle
            w1, n1 = self._second
            self._second = None
            j = self._j - 1
            return self.src.step_word2(self.g, self.tick, j, w1, n1)
        j = self._j
        self._j += 1
        w = self.src.step_word(self.g, self.tick, j)
        # A randint(100) from move_obstacles :332 is followed by randint(9) iff >= threshold.
        # n == 100 is only ever drawn there (goal lists of length 101 are rejected by the harness).
        if n == 100 and D.mulhi(w, 100) >= self.env.obs_uncertainity_threshold:
            self._second = (w, 100)
        return w


# --------------------------------------------------------------------------- loading
_loaded = {}


def load_ballenv_module():
    """Import the reference's gym env module, unedited."""
    if "be" not in _loaded:
        if not reference_available():
            raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
        _install_gym_standin()
        if REFERENCE_ROOT not in sys.path:
            sys.path.insert(0, REFERENCE_ROOT)
        import gym_ballenv.envs.ballenv_env as be      # noqa: E402  (the reference, not this repo)
        _loaded["be"] = be
    return _loaded["be"]


def load_prep_state(env, names=("prep_state2", "prep_state4")):
    """AST-lift the observation builders of examples/ball_cnn_ac3.py and bind
    them to ``env`` (they read the script globals ``env`` and ``device``)."""
    import torch
    path = os.path.join(REFERENCE_ROOT, "examples", "ball_cnn_ac3.py")
    with open(path) as f:
        tree = ast.parse(f.read(), path)
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    ns = {"np": _np, "torch": torch, "math": math, "env": env, "device": torch.device("cpu")}
    exec(compile(ast.Module(body=keep, type_ignores=[]), path, "exec"), ns)
    return tuple(ns[n] for n in names)


def load_a2c(window, gamma=0.99):
    """AST-lift ``Policy`` (examples/ball_cnn_ac3.py:109-146) and ``finish_episode`` (:222-246) of the reference's
    actor-critic script (the module itself cannot be imported: argparse, gym.make, matplotlib and readchar at top level)
    into a namespace that provides what they read as script globals.  -> namespace with ``policy`` (a fresh
    ``Policy(window)`` on the CPU), ``finish_episode``, ``SavedAction``, ``printed`` (what finish_episode printed: the
    loss) and an ``optimizer`` whose step changes nothing (lr = 0), so that the gradients of the call stay readable."""
    import collections
    import torch
    import torch.nn as nn
    import torch.nn.functional as F
    path = os.path.join(REFERENCE_ROOT, "examples", "ball_cnn_ac3.py")
    with open(path) as f:
        tree = ast.parse(f.read(), path)
    keep = [n for n in tree.body
            if (isinstance(n, ast.ClassDef) and n.name == "Policy") or (isinstance(n, ast.FunctionDef) and n.name == "finish_episode")]
    assert len(keep) == 2
    printed = []
    ns = {"np": _np, "torch": torch, "nn": nn, "F": F, "math": math, "device": torch.device("cpu"),
          "eps": _np.finfo(_np.float32).eps.item(),                      # :76
          "args": type("Args", (), {"gamma": gamma})(),
          "SavedAction": collections.namedtuple("SavedAction", ["log_prob", "value"]),   # :78
          "print": lambda *a, **k: printed.append(a), "printed": printed}
    exec(compile(ast.Module(body=keep, type_ignores=[]), path, "exec"), ns)
    ns["policy"] = ns["Policy"](window)
    ns["optimizer"] = torch.optim.SGD(ns["policy"].parameters(), lr=0.0)
    return ns


REFERENCE_CHECKPOINT_W5 = os.path.join(REFERENCE_ROOT, "examples", "stored_models", "ball_state3",
                                       "2layer+dropout+randpos", "episode_2500.pth")   # the README's WINDOW = 5 run


def load_legacy_block_state():
    """AST-lift the legacy 29-float observation of examples/ball_env_reinforce.py:130-172 (prep_state2 + block_to_arrpos).
    The script is Python 2: with integer coordinates ``x_dist/abs(x_dist)`` is an integer there and the array index
    ``4+pos`` an int; under Python 3 the same expressions are floats and numpy refuses the index.  The lifted
    ``block_to_arrpos`` is therefore wrapped in ``int()`` (the value is integral in both) - nothing else is touched."""
    if "legacy" not in _loaded:
        path = os.path.join(REFERENCE_ROOT, "examples", "ball_env_reinforce.py")
        with open(path) as f:
            tree = ast.parse(f.read(), path)
        keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("prep_state2", "block_to_arrpos")]
        ns = {"np": _np, "math": math}
        exec(compile(ast.Module(body=keep, type_ignores=[]), path, "exec"), ns)
        raw = ns["block_to_arrpos"]
        ns["block_to_arrpos"] = lambda x, y: int(raw(x, y))
        _loaded["legacy"] = ns["prep_state2"]
    return _loaded["legacy"]


def default_args(**over):
    """Defaults of examples/ball_cnn_ac3.py:40-51 as the Namespace the
    reference's customize_environment (ballenv_env.py:87-109) consumes."""
    a = dict(static_obstacles=13, dynamic_obstacles=5, obstacle_speed=[1, 1, 1, 1, 1],
             obs_goal_position=['12,122', '123,93', '87,150', '430,440', '230,11'],
             time_step_for_change=50, rd_th_obs=60, rd_th_agent=80,
             static_thresholds=[0, 0], dynamic_thresholds=[10, 10],
             static_penalty=[1, 1], dynamic_penalty=[4000, 8000])
    a.update(over)
    return Namespace(**a)


class ReferenceEnv(object):
    """One reference ``BallEnv`` driven through a draw router."""

    def __init__(self, args, router):
        be = load_ballenv_module()
        self.be = be
        self.router = router
        be.np = _NumpyProxy(router)            # step 3 of the recipe (module global, not a file edit)
        router.context = "ctor"
        self.env = be.BallEnv()
        if hasattr(router, "env"):
            router.env = self.env
        self.env.customize_environment(args)
        self.prep_state2, self.prep_state4 = load_prep_state(self.env)

    def _bind(self):
        self.be.np = _NumpyProxy(self.router)

    def reset(self):
        self._bind()
        self.env.static_obstacle_list = []      # step 5 of the recipe
        self.env.dynamic_obstacle_list = []
        if hasattr(self.router, "begin_reset"):
            self.router.begin_reset()
        else:
            self.router.context = "reset"
        return self.env.reset()

    def step(self, action):
        self._bind()
        if hasattr(self.router, "begin_step"):
            self.router.begin_step()
        else:
            self.router.context = "step"
        out = self.env.step(action)
        if hasattr(self.router, "end_step"):
            self.router.end_step()
        return out

    def inject(self, agent, goal, dist, obstacles, dyn_goal_idx=None, dyn_counter=None,
               old_dist=None, total_distance=None, acc=0.0):
        """Overwrite the env state (positions of already-created obstacles)."""
        env = self.env
        assert len(obstacles) == len(env.obstacle_list)
        for o, (x, y) in zip(env.obstacle_list, obstacles):
            o.x, o.y = x, y
        ks = env.no_of_static_obstacles
        for j, o in enumerate(env.obstacle_list[ks:]):
            if dyn_goal_idx is not None:
                o.curr_goal = env.obstacle_goal_list[dyn_goal_idx[j]]
            if dyn_counter is not None:
                o.curr_counter = dyn_counter[j]
        env.goal_x, env.goal_y = goal
        env.state = [tuple(agent), tuple(goal), dist] + [tuple(p) for p in obstacles]
        env.old_dist = dist if old_dist is None else old_dist
        if total_distance is not None:
            env.total_distance = total_distance
        env.total_reward_accumulated = acc

    def observe(self, state, window):
        """prep_state4 of the reference -> numpy float32 [4 + W*W]."""
        return self.prep_state4(state, window).numpy().reshape(-1)


# --------------------------------------------------------------------------- pygame ruleset (ballenv_pygame.py)
class PygameRouter(object):
    """Answers ``createBoard``'s draws (ballenv_pygame.py:454-457,468-482,489-498 and Obstacle.__init__ :24-33)
    from an addressed draw source.  reset: ranf x2 (goal), ranf x2 (agent), ranf x2 per agent redraw, then per
    static obstacle a randint (x, y) pair per attempt; ``step`` draws nothing."""

    def __init__(self, source, g):
        self.src = source
        self.g = g
        self.env = None
        self.context = "ctor"
        self.episode = -1
        self.trace = []

    def begin_reset(self):
        self.context = "reset"
        self.episode += 1
        self._nf = 0
        self._pair = None
        self._last_len = -1
        self._attempt = 0

    def begin_reset_fixed(self, goal=(145, 120)):
        """resetFixedstate (ballenv_pygame.py:589-624) draws nothing but agent positions, two ranf each: the first of an
        outer attempt, then a redraw while the agent is closer than 50 to the (fixed) goal.  Which of the two the next
        pair is follows from the pair just handed out, so the router can address it as (outer, inner)."""
        self.context = "reset_fixed"
        self.episode += 1
        self._fixed_goal = goal
        self._outer, self._inner = -1, 0
        self._half = None          # words of the current pair while its second ranf is pending
        self._last = None          # the agent position handed out last

    def _ranf_fixed(self):
        if self._half is None:
            if self._last is not None and math.sqrt(math.pow(self._fixed_goal[0] - self._last[0], 2) +
                                                    math.pow(self._fixed_goal[1] - self._last[1], 2)) < 50:
                self._inner += 1
            else:
                self._outer, self._inner = self._outer + 1, 0
            self._half = self.src.reset_words(self.g, self.episode, D.RK_FIXED_AGENT, item=self._outer,
                                              attempt=self._inner, count=4)
            v = D.ranf_from_words(self._half[0], self._half[1])
            self._x = 0 + v * (100 - 0)
        else:
            v = D.ranf_from_words(self._half[2], self._half[3])
            self._last = (self._x, 0 + v * (100 - 0))
            self._half = None
        self.trace.append((self.context, 0.0, 1.0, v))
        return v

    def ranf(self):
        if self.context == "reset_fixed":
            return self._ranf_fixed()
        assert self.context == "reset"
        k = self._nf
        self._nf += 1
        if k < 4:      # goal (block 0: words 0-1, 2-3), first agent draw (block 1)
            w = self.src.reset_words(self.g, self.episode, D.RK_HEAD, item=k // 2, count=4)
        else:          # agent redraws, two ranf per attempt
            w = self.src.reset_words(self.g, self.episode, D.RK_AGENT_REDRAW, attempt=(k - 4) // 2, count=4)
        v = D.ranf_from_words(w[2 * (k % 2)], w[2 * (k % 2) + 1])
        self.trace.append((self.context, 0.0, 1.0, v))
        return v

    def randint(self, low, high):
        n = high - low
        assert self.context == "reset"
        if self._pair is None:
            cur = len(self.env.obstacle_list)
            if cur != self._last_len:
                self._last_len, self._attempt = cur, 0
            else:
                self._attempt += 1
            self._pair = self.src.reset_words(self.g, self.episode, D.RK_STATIC, cur, self._attempt)
            w = self._pair[0]
        else:
            w = self._pair[1]
            self._pair = None
        v = low + D.mulhi(w, n)
        self.trace.append((self.context, low, high, v))
        return v


def _install_pygame_standins():
    """createBoard(display=False) only touches pygame.init / pygame.time.Clock; its step() calls
    featureExtractor.featureExtractor, which needs CUDA and does not influence state, reward or done."""
    import types
    if "pygame" not in sys.modules:
        pg = types.ModuleType("pygame")
        pg.init = lambda *a, **k: (0, 0)
        pg.quit = lambda *a, **k: None
        pg.time = types.SimpleNamespace(Clock=lambda: types.SimpleNamespace(tick=lambda *a, **k: None))
        sys.modules["pygame"] = pg
    if "featureExtractor" not in sys.modules:
        fe = types.ModuleType("featureExtractor")
        fe.featureExtractor = lambda *a, **k: None
        sys.modules["featureExtractor"] = fe


def load_pygame_module():
    if "pg" not in _loaded:
        import contextlib
        import importlib.util
        import io
        _install_pygame_standins()
        spec = importlib.util.spec_from_file_location("ref_ballenv_pygame", os.path.join(REFERENCE_ROOT, "ballenv_pygame.py"))
        mod = importlib.util.module_from_spec(spec)
        with contextlib.redirect_stdout(io.StringIO()):
            spec.loader.exec_module(mod)
        _loaded["pg"] = mod
    return _loaded["pg"]


class ReferencePygameEnv(object):
    """One reference ``createBoard`` (ballenv_pygame.py:314-706), unedited, driven through a draw router."""

    def __init__(self, static_obstacles, router, agent_radius=10, static_obstacle_radius=10):
        import contextlib
        import io
        self.pg = load_pygame_module()
        self.router = router
        self.pg.np = _NumpyProxy(router)
        with contextlib.redirect_stdout(io.StringIO()):      # the ctor prints pygame.init() and "here"
            self.env = self.pg.createBoard(display=False, static_obstacles=static_obstacles, agent_radius=agent_radius,
                                           static_obstacle_radius=static_obstacle_radius)
        router.env = self.env

    def reset(self):
        self.pg.np = _NumpyProxy(self.router)
        self.env.static_obstacle_list = []
        self.router.begin_reset()
        return self.env.reset()

    def reset_fixed(self):
        self.pg.np = _NumpyProxy(self.router)
        self.router.begin_reset_fixed()
        return self.env.resetFixedstate()

    def step(self, action):
        self.pg.np = _NumpyProxy(self.router)
        self.router.context = "step"
        return self.env.step(action)


def reference_features(agent, goal, obstacles, agent_rad=10, obstacle_rad=20, agent_vel=(0, 0), obstacle_vel=(0, 0)):
    """The reference's own numpy helpers (featureExtractor.py:36-193), concatenated exactly as
    featureExtractor() does (:249-257) - only its final ``.to(cuda)`` (:262-264) is left out."""
    if "fe" not in _loaded:
        import importlib.util
        if "pygame" not in sys.modules or not hasattr(sys.modules["pygame"], "init"):
            _install_pygame_standins()
        real = sys.modules.pop("featureExtractor", None)      # the stand-in used by the createBoard driver
        spec = importlib.util.spec_from_file_location("ref_featureExtractor", os.path.join(REFERENCE_ROOT, "featureExtractor.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        if real is not None:
            sys.modules["featureExtractor"] = real
        _loaded["fe"] = mod
    fe = _loaded["fe"]
    obs = [types.SimpleNamespace(x=o[0], y=o[1], rad=obstacle_rad, vel_x=obstacle_vel[0], vel_y=obstacle_vel[1])
           for o in obstacles]
    state = [tuple(agent), tuple(goal), 0.0, False] + [tuple(o) for o in obstacles]
    parts = (fe.calcDistanceFromGoal(state, 5), fe.relativeGoalPos(state), fe.densityFeatures(state, obs, agent_rad, 10),
             fe.speedOrientationFeatures(state, obs, agent_rad, agent_vel).reshape(9),
             fe.socialForcesFeatures(state, obs, agent_rad, agent_vel))
    return _np.concatenate(parts).astype(_np.float64)
