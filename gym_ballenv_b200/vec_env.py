"""N-environment vector wrapper over the fused CUDA step+observe kernel.

New relative to the reference (which steps one Python env per process): ``BallVecEnv`` advances
N independent ball environments per launch and returns torch CUDA tensors.  Semantics of a single
environment are those of the reference's ``BallEnv`` (gym_ballenv/envs/ballenv_env.py:37-289) or,
with ``ruleset="pygame"``, ``createBoard`` (ballenv_pygame.py:314-706); the observation is
``prep_state4(state, WINDOW)`` of examples/ball_cnn_ac3.py:384-412.  Added: the TimeLimit(1000)
of the gym registration (gym_ballenv/__init__.py:7) as a ``truncated`` flag OR-ed into ``done``, and
auto-reset of finished environments inside the same launch (the returned observation is the first
observation of the new episode; reward/done belong to the finished one).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from . import _lib as L
from ._lib import LIB, check
from .config import EnvConfig

# agent action table of the training loops (examples/ball_cnn_ac3.py:530)
MOVE_LIST = [(1, 1), (1, -1), (1, 0), (0, 1), (0, -1), (0, 0), (-1, 1), (-1, 0), (-1, -1)]

_RULESETS = {"gym": L.RULESET_GYM, "pygame": L.RULESET_PYGAME}
_OBS_FORMATS = {torch.float32: L.OBS_F32, torch.uint8: L.OBS_U8, "bits": L.OBS_BITS}


class BallVecEnv:
    """N ball environments resident on one GPU.

    step(actions) accepts
      * int64 / int32 / uint8 ``[N]`` indices into ``MOVE_LIST``, or
      * float32 / float64 ``[N, 2]`` raw ``(dx, dy)`` (what ``BallEnv.step`` indexes, ballenv_env.py:247-248)
    and returns ``(obs [N, 4 + W*W], reward [N], done [N] bool, info)``.  The returned tensors are views of
    env-owned double buffers: they stay valid until the step after next (``clone()`` to keep them).
    ``info["state"]`` / ``state_views`` are live views of the state arrays: edit them through ``set_state`` or call
    ``state_written()`` afterwards.
    """

    def __init__(self, num_envs: int, window: int = 5, config: Optional[EnvConfig] = None, ruleset: str = "gym",
                 device="cuda", seed: int = 0, obs_dtype=torch.float32, parity: bool = False,
                 auto_reset: bool = True, max_episode_steps: Optional[int] = None, global_env_offset: int = 0):
        if not torch.cuda.is_available():
            raise RuntimeError("gym_ballenv_b200 needs a CUDA device (there is no CPU fallback)")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("device must be a CUDA device")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.num_envs = int(num_envs)
        self.window = int(window)
        self.ruleset = ruleset
        if config is None:
            config = EnvConfig() if ruleset == "gym" else EnvConfig.pygame_default()
        self.config = config
        self.parity = bool(parity)
        self.seed = int(seed)
        self.global_env_offset = int(global_env_offset)
        if max_episode_steps is None:
            max_episode_steps = 1000 if ruleset == "gym" else 0
        self.max_episode_steps = int(max_episode_steps)
        self.auto_reset = bool(auto_reset)
        self.obs_dtype = obs_dtype
        self._ccfg = config.to_c(window, _RULESETS[ruleset], L.F64 if parity else L.F32, _OBS_FORMATS[obs_dtype],
                                 self.max_episode_steps, self.auto_reset)
        nbytes = check(LIB.ballenv_state_bytes(C.byref(self._ccfg), self.num_envs))
        with torch.cuda.device(self.device):
            self._arena = torch.zeros(nbytes, dtype=torch.uint8, device=self.device)
            torch.cuda.synchronize(self.device)
        h = C.c_void_p()
        check(LIB.ballenv_create(C.byref(self._ccfg), self.num_envs, self.global_env_offset, self.device.index,
                                 C.c_uint64(self.seed & (2 ** 64 - 1)), C.c_void_p(self._arena.data_ptr()), C.byref(h)))
        self._h = h
        self._ptrs = L.BallenvStatePtrs()
        check(LIB.ballenv_state_ptrs(self._h, C.byref(self._ptrs)))
        # positions / rewards are float64 in parity mode - and always for the pygame ruleset (non-integral coordinates:
        # fp32 storage cannot hold the 1e-5 reward tolerance there; the library promotes it, real_bytes tells)
        self._real = torch.float64 if int(self._ptrs.real_bytes) == 8 else torch.float32
        self.parity = self._real == torch.float64
        self._make_views()
        self.obs_row = int(self._ptrs.obs_row_elems)
        obs_t = torch.int32 if obs_dtype == "bits" else obs_dtype
        n = self.num_envs
        self._bufs = [dict(obs=torch.empty((n, self.obs_row), dtype=obs_t, device=self.device),
                           reward=torch.empty(n, dtype=self._real, device=self.device),
                           done=torch.empty(n, dtype=torch.uint8, device=self.device)) for _ in range(2)]
        self._flip = 0
        self._tape_refs = None
        # what step() hands back, built once: a launch of one step is a few microseconds, the Python around it must
        # not cost more (tools/step_overhead.py) - the buffers' pointers, the bool views of `done`, the info dict
        self._step_out = [(C.c_void_p(b["obs"].data_ptr()), C.c_void_p(b["reward"].data_ptr()), C.c_void_p(b["done"].data_ptr()),
                           b["obs"], b["reward"], b["done"].view(torch.bool)) for b in self._bufs]
        self._info = {"flags": self.state_views["flags"], "state": self.state_views}
        self._dev_index = self.device.index
        self._raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)

    # ------------------------------------------------------------------ plumbing
    def _stream(self):
        if self._raw_stream is not None:      # the stream's handle without building a torch.cuda.Stream object
            return C.c_void_p(self._raw_stream(self._dev_index))
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _views_of(self, arena: torch.Tensor) -> Dict[str, torch.Tensor]:
        """Typed views of the SoA state inside ``arena`` (the device arena, or a host copy of it).
        Obstacle fields are stored environment-major [n_stride][K padded to 4] (include/ballenv.h) and are
        exposed transposed as [K, N] like the reference's state list order."""
        p, S, n = self._ptrs, int(self._ptrs.n_stride), self.num_envs
        base = self._arena.data_ptr()
        rb = int(p.real_bytes)
        ks, kd = self.config.static_obstacles, self.config.dynamic_obstacles
        ss, ds = int(p.static_stride), int(p.dynamic_stride)

        def view(ptr, nbytes, dtype, shape):
            off = ptr - base
            return arena[off:off + nbytes].view(dtype).view(shape)

        v = {}
        for name in ("agent_x", "agent_y", "goal_x", "goal_y"):
            v[name] = view(getattr(p, name), rb * S, self._real, (S,))[:n]
        for name in ("dist", "total_distance", "acc_reward"):
            v[name] = view(getattr(p, name), 8 * S, torch.float64, (S,))[:n]
        for name in ("ep_len", "episode", "tick"):
            v[name] = view(getattr(p, name), 4 * S, torch.int32, (S,))[:n]
        for name, k, st in (("static_x", ks, ss), ("static_y", ks, ss), ("dynamic_x", kd, ds), ("dynamic_y", kd, ds)):
            v[name] = (view(getattr(p, name), rb * S * st, self._real, (S, st))[:n, :k].t() if k
                       else torch.empty((0, n), dtype=self._real, device=arena.device))
        v["dynamic_meta"] = (view(p.dynamic_meta, 4 * S * ds, torch.int32, (S, ds))[:n, :kd].t() if kd
                             else torch.empty((0, n), dtype=torch.int32, device=arena.device))
        v["flags"] = view(p.flags, S, torch.uint8, (S,))[:n]
        v["stats"] = view(p.stats, 8 * L.NUM_STATS, torch.float64, (L.NUM_STATS,))
        return v

    def _make_views(self):
        self.state_views: Dict[str, torch.Tensor] = self._views_of(self._arena)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            LIB.ballenv_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ API
    @property
    def observation_size(self):
        return 4 + self.window * self.window

    def reset(self, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Reset all environments (or those where ``mask`` is non-zero) -> observation [N, 4 + W*W]."""
        buf = self._next_buf()
        mptr = None
        if mask is not None:
            mask = mask.to(device=self.device, dtype=torch.uint8).contiguous()
            assert mask.shape == (self.num_envs,)
            mptr = C.c_void_p(mask.data_ptr())
        check(LIB.ballenv_reset(self._h, mptr, C.c_void_p(buf["obs"].data_ptr()), self._stream()))
        return buf["obs"]

    def reset_fixed(self, goal=(145.0, 120.0), mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        """createBoard.resetFixedstate (ballenv_pygame.py:589-624; pygame ruleset): new episodes with the goal at
        ``goal`` and the obstacles kept -> observation [N, 4 + W*W]."""
        buf = self._next_buf()
        mptr = None
        if mask is not None:
            mask = mask.to(device=self.device, dtype=torch.uint8).contiguous()
            assert mask.shape == (self.num_envs,)
            mptr = C.c_void_p(mask.data_ptr())
        check(LIB.ballenv_reset_fixed(self._h, mptr, float(goal[0]), float(goal[1]), C.c_void_p(buf["obs"].data_ptr()),
                                      self._stream()))
        return buf["obs"]

    def observe(self) -> torch.Tensor:
        """prep_state4 of the current state, without stepping."""
        buf = self._next_buf()
        check(LIB.ballenv_observe(self._h, C.c_void_p(buf["obs"].data_ptr()), self._stream()))
        return buf["obs"]

    def sensor_readings(self) -> torch.Tensor:
        """The 20-float social-navigation features of the current state (featureExtractor.py:247-265, what
        createBoard keeps in ``self.sensor_readings``): float32 [N, 20] on the device."""
        out = torch.empty((self.num_envs, 20), dtype=torch.float32, device=self.device)
        check(LIB.ballenv_observe_features(self._h, C.c_void_p(out.data_ptr()), self._stream()))
        return out

    def block_counts(self) -> torch.Tensor:
        """The legacy 29-float observation of the REINFORCE / imitation scripts (``prep_state2`` of
        examples/ball_env_reinforce.py:130-172): 4 goal-quadrant bits + a 5 x 5 grid of obstacle counts in 20-pixel
        blocks around the agent.  float32 [N, 29] on the device, of the current state."""
        out = torch.empty((self.num_envs, 29), dtype=torch.float32, device=self.device)
        check(LIB.ballenv_observe_blocks(self._h, C.c_void_p(out.data_ptr()), self._stream()))
        return out

    def rgb_patches(self, width: int = 100, size: int = 40, interp: str = "bicubic", dtype=torch.float32,
                    out: torch.Tensor = None) -> torch.Tensor:
        """The pixel policies' observation (``extract_patch`` of examples/ball_cnn_reinforce.py:120-163 and
        examples/ball_cnn_ac3.py:248-307): the viewer's frame (ballenv_env.py:357-386) cropped to ``width`` x ``width``
        around the agent and resized to ``size`` x ``size`` as PIL does (``interp`` "bicubic" or "bilinear"), for every
        environment at once: float32 [N, 3, size, size] in [0, 1] (ToTensor) or uint8, on the device, of the current
        state.  Rendered on the GPU - no pyglet, no host round trip."""
        if interp not in ("bilinear", "bicubic"):
            raise ValueError("interp must be 'bilinear' or 'bicubic'")
        if dtype not in (torch.float32, torch.uint8):
            raise ValueError("patches are float32 or uint8")
        if out is None:
            out = torch.empty((self.num_envs, 3, size, size), dtype=dtype, device=self.device)
        elif out.shape != (self.num_envs, 3, size, size) or out.dtype != dtype or not out.is_contiguous() or out.device != self.device:
            raise ValueError("out must be a contiguous [N, 3, size, size] tensor of the requested dtype on the env's device")
        check(LIB.ballenv_observe_patches(self._h, C.c_void_p(out.data_ptr()), int(width), int(size),
                                          1 if interp == "bicubic" else 0,
                                          L.OBS_F32 if dtype == torch.float32 else L.OBS_U8, self._stream()))
        return out

    def _next_buf(self):
        self._flip ^= 1
        return self._bufs[self._flip]

    _INDEX_KINDS = {torch.int64: L.ACT_INDEX_I64, torch.int32: L.ACT_INDEX_I32, torch.uint8: L.ACT_INDEX_U8}
    _XY_KINDS = {torch.float32: L.ACT_XY_F32, torch.float64: L.ACT_XY_F64}

    @classmethod
    def _action_kind(cls, actions: torch.Tensor, n: int):
        if actions.dim() == 1:
            kind = cls._INDEX_KINDS.get(actions.dtype)
            if kind is None or actions.shape[0] != n:
                raise ValueError("index actions must be int64/int32/uint8 [N]")
            return kind
        if actions.dim() == 2 and actions.shape == (n, 2):
            kind = cls._XY_KINDS.get(actions.dtype)
            if kind is None:
                raise ValueError("raw actions must be float32/float64 [N, 2]")
            return kind
        raise ValueError("actions must be [N] indices or [N, 2] raw (dx, dy)")

    def step(self, actions: torch.Tensor):
        """-> (obs, reward, done, info): views of env-owned double buffers, valid until the step after next; ``info`` is
        the same dict on every call (flags and live state views)."""
        if actions.device != self.device:
            raise ValueError("actions must live on %s" % self.device)
        kind = self._action_kind(actions, self.num_envs)
        if not actions.is_contiguous():
            actions = actions.contiguous()
        self._flip ^= 1
        p_obs, p_rew, p_done, obs, reward, done = self._step_out[self._flip]
        rc = LIB.ballenv_step(self._h, actions.data_ptr(), kind, p_obs, p_rew, p_done, self._stream())
        if rc < 0:
            check(rc)
        return obs, reward, done, self._info

    def step_into(self, actions: torch.Tensor, obs_out: torch.Tensor, reward_out: torch.Tensor, done_out: torch.Tensor):
        """step() writing into caller-owned device tensors (obs [N, row], reward [N], done [N] uint8): fixed
        addresses, which is what CUDA-graph capture of a policy-in-the-loop rollout needs."""
        kind = self._action_kind(actions, self.num_envs)
        if (actions.device != self.device or not actions.is_contiguous() or tuple(obs_out.shape) != (self.num_envs, self.obs_row)
                or obs_out.dtype != self._bufs[0]["obs"].dtype or reward_out.dtype != self._real
                or done_out.dtype != torch.uint8 or reward_out.numel() != self.num_envs or done_out.numel() != self.num_envs
                or not (obs_out.is_contiguous() and reward_out.is_contiguous() and done_out.is_contiguous())):
            raise ValueError("step_into needs contiguous device tensors of the shapes / dtypes step() returns")
        check(LIB.ballenv_step(self._h, C.c_void_p(actions.data_ptr()), kind, C.c_void_p(obs_out.data_ptr()),
                               C.c_void_p(reward_out.data_ptr()), C.c_void_p(done_out.data_ptr()), self._stream()))

    def rollout_policy(self, policy, n_steps: int, first_obs: torch.Tensor, obs_out: torch.Tensor,
                       actions_out: torch.Tensor, reward_out: torch.Tensor, done_out: torch.Tensor, greedy: bool = False,
                       policy_out: Optional[torch.Tensor] = None):
        """n_steps of the loop of examples/ball_cnn_ac3.py:553-613 - observe, ``Policy`` forward, Categorical sample,
        step - for all environments in ONE launch: the kernel evaluates the MLP itself between two steps
        (ballenv_rollout_policy).  ``policy``: a module with ``fc1`` / ``action_head`` Linear layers (a2c.Policy; the
        value head is not needed to act).  first_obs [N, row] float32: the current observation; obs_out [T, N, row],
        actions_out [T, N] int64, reward_out [T, N] float32, done_out [T, N] uint8 are written; ``policy_out`` (optional,
        float32 [T, N, 10]) receives the 9 action probabilities and the value of every forward pass (``policy`` then
        needs a ``value_head``): what the update would otherwise recompute.  Raises BallenvError
        for configurations without such a kernel (WINDOW other than 5 / 10, more than 64 obstacles, parity mode)."""
        n, row, dev = self.num_envs, self.obs_row, self.device
        w1, b1 = policy.fc1.weight, policy.fc1.bias
        w2, b2 = policy.action_head.weight, policy.action_head.bias
        for t in (w1, b1, w2, b2):
            if t.dtype != torch.float32 or t.device != dev or not t.is_contiguous():
                raise ValueError("policy parameters must be contiguous float32 tensors on the env's device")
        hidden = w1.shape[0]
        if tuple(w1.shape) != (hidden, row) or tuple(w2.shape) != (9, hidden):
            raise ValueError("policy must map %d inputs -> hidden -> 9 actions" % row)
        def ok(t, shape, dtype):
            return tuple(t.shape) == shape and t.dtype == dtype and t.device == dev and t.is_contiguous()
        if not (ok(first_obs, (n, row), torch.float32) and ok(obs_out, (n_steps, n, row), torch.float32)
                and ok(actions_out, (n_steps, n), torch.int64) and ok(reward_out, (n_steps, n), torch.float32)
                and ok(done_out, (n_steps, n), torch.uint8)):
            raise ValueError("rollout_policy needs contiguous device tensors: first_obs [N, row] f32, obs_out [T, N, row] f32, "
                             "actions_out [T, N] i64, reward_out [T, N] f32, done_out [T, N] u8")
        pol = L.BallenvPolicyMLP(n_inputs=row, hidden=hidden, greedy=1 if greedy else 0, reserved=0,
                                 fc1_weight=w1.data_ptr(), fc1_bias=b1.data_ptr(), action_weight=w2.data_ptr(),
                                 action_bias=b2.data_ptr())
        po = None
        if policy_out is not None:
            wv, bv = policy.value_head.weight, policy.value_head.bias
            if not (ok(policy_out, (n_steps, n, 10), torch.float32) and ok(wv, (1, hidden), torch.float32) and ok(bv, (1,), torch.float32)):
                raise ValueError("policy_out must be a contiguous float32 [T, N, 10] device tensor (and the policy needs a value_head)")
            pol.value_weight, pol.value_bias = wv.data_ptr(), bv.data_ptr()
            po = C.c_void_p(policy_out.data_ptr())
        check(LIB.ballenv_rollout_policy(self._h, C.byref(pol), int(n_steps), C.c_void_p(first_obs.data_ptr()),
                                         C.c_void_p(obs_out.data_ptr()), C.c_void_p(actions_out.data_ptr()),
                                         C.c_void_p(reward_out.data_ptr()), C.c_void_p(done_out.data_ptr()), po, self._stream()))

    def alloc_rollout(self, T: int, keep_all_obs: bool = False):
        """Rollout buffers for step_many(out=...): (obs [T, N, row] or [N, row], reward [T, N], done [T, N] uint8)."""
        n = self.num_envs
        obs = torch.empty(((T, n, self.obs_row) if keep_all_obs else (n, self.obs_row)),
                          dtype=self._bufs[0]["obs"].dtype, device=self.device)
        reward = torch.empty((T, n), dtype=self._real, device=self.device)
        done = torch.empty((T, n), dtype=torch.uint8, device=self.device)
        return obs, reward, done

    def step_many(self, actions: torch.Tensor, keep_all_obs: bool = False, out=None):
        """T steps in one call; actions [T, N] indices or [T, N, 2] raw.
        -> (obs [T, N, row] or [N, row], reward [T, N], done [T, N] bool).  ``out`` = alloc_rollout(T, ...)
        buffers to write into (a training loop's rollout storage) instead of fresh tensors."""
        T = actions.shape[0]
        kind = self._action_kind(actions[0], self.num_envs)
        if actions.device != self.device:
            raise ValueError("actions must live on %s" % self.device)
        actions = actions.contiguous()
        obs, reward, done = out if out is not None else self.alloc_rollout(T, keep_all_obs)
        if out is not None:
            want = (T, self.num_envs, self.obs_row) if keep_all_obs else (self.num_envs, self.obs_row)
            if (tuple(obs.shape) != want or tuple(reward.shape) != (T, self.num_envs)
                    or tuple(done.shape) != (T, self.num_envs) or done.dtype != torch.uint8
                    or not (obs.is_contiguous() and reward.is_contiguous() and done.is_contiguous())):
                raise ValueError("out buffers do not match alloc_rollout(%d, %s)" % (T, keep_all_obs))
        check(LIB.ballenv_step_many(self._h, C.c_void_p(actions.data_ptr()), kind, T, C.c_void_p(obs.data_ptr()),
                                    1 if keep_all_obs else 0, C.c_void_p(reward.data_ptr()),
                                    C.c_void_p(done.data_ptr()), self._stream()))
        return obs, reward, done.view(torch.bool)

    def step_host(self, actions, obs_out, reward_out, done_out):
        """Host-buffer step through the C ABI (ballenv_step_host): ``actions`` and the three outputs are CPU
        tensors (pinned for speed); copies both ways happen inside the call, which returns synchronised."""
        kind = self._action_kind(actions, self.num_envs)
        check(LIB.ballenv_step_host(self._h, C.c_void_p(actions.data_ptr()), kind, C.c_void_p(obs_out.data_ptr()),
                                    C.c_void_p(reward_out.data_ptr()), C.c_void_p(done_out.data_ptr()),
                                    self._stream()))
        return obs_out, reward_out, done_out

    def step_many_host(self, actions, obs_out, reward_out, done_out):
        """T host-buffer steps as one pipelined call (ballenv_step_many_host): ``actions`` [T, N] and the outputs
        ``obs_out`` [T, N, row], ``reward_out`` [T, N], ``done_out`` [T, N] uint8 are CPU tensors (pinned for speed).
        Copies in, kernels and copies out of consecutive steps overlap; the call returns synchronised."""
        T = int(actions.shape[0])
        kind = self._action_kind(actions[0], self.num_envs)
        n = self.num_envs
        if (not actions.is_contiguous() or tuple(obs_out.shape) != (T, n, self.obs_row) or tuple(reward_out.shape) != (T, n)
                or tuple(done_out.shape) != (T, n) or done_out.dtype != torch.uint8 or reward_out.dtype != self._real
                or obs_out.dtype != self._bufs[0]["obs"].dtype
                or not (obs_out.is_contiguous() and reward_out.is_contiguous() and done_out.is_contiguous())
                or any(t.device.type != "cpu" for t in (actions, obs_out, reward_out, done_out))):
            raise ValueError("step_many_host needs contiguous CPU tensors: actions [T, N], obs [T, N, row], reward [T, N], "
                             "done [T, N] uint8")
        check(LIB.ballenv_step_many_host(self._h, C.c_void_p(actions.data_ptr()), kind, T, C.c_void_p(obs_out.data_ptr()),
                                         C.c_void_p(reward_out.data_ptr()), C.c_void_p(done_out.data_ptr()),
                                         self._stream()))
        return obs_out, reward_out, done_out

    # ------------------------------------------------------------------ state injection / inspection
    def get_state(self) -> Dict[str, torch.Tensor]:
        """Copies of the SoA state.  Obstacles are [K, N]; ``dynamic_goal`` / ``dynamic_counter`` unpack dynamic_meta."""
        out = {k: v.clone() for k, v in self.state_views.items() if k != "stats"}
        meta = out.pop("dynamic_meta")
        out["dynamic_goal"] = meta & 0xff
        out["dynamic_counter"] = meta >> 8
        return out

    def set_state(self, **fields):
        """Overwrite state arrays (same names/shapes as get_state()).  Used for parity injection."""
        fields = dict(fields)
        if "dynamic_goal" in fields or "dynamic_counter" in fields:
            meta = self.state_views["dynamic_meta"]
            goal = fields.pop("dynamic_goal", meta & 0xff)
            cnt = fields.pop("dynamic_counter", meta >> 8)
            goal = torch.as_tensor(goal, device=self.device).to(torch.int32)
            cnt = torch.as_tensor(cnt, device=self.device).to(torch.int32)
            meta.copy_(goal | (cnt << 8))
        for k, val in fields.items():
            dst = self.state_views[k]
            dst.copy_(torch.as_tensor(val, device=self.device).to(dst.dtype))
        self.state_written()

    def state_written(self):
        """Call after writing into ``state_views`` / ``info["state"]`` directly (set_state does it itself): the library
        re-validates what its production kernels assume about the coordinates (ballenv_state_written)."""
        check(LIB.ballenv_state_written(self._h, self._stream()))

    def set_draw_tape(self, step_tape=None, reset_tape=None, attempts: int = 1):
        """Parity mode: inject the random words (uint32, held as int64/uint32 CPU tensors or numpy arrays).
        step_tape [T, N, Kd, 2]; reset_tape [E, N, 4 + 2*attempts*Ks + 2*Kd]; None clears."""
        import numpy as np
        st = rt = None
        sp = rp = None
        T = E = 0
        if step_tape is not None:
            st = np.ascontiguousarray(np.asarray(step_tape), dtype=np.uint32)
            assert st.shape[1:] == (self.num_envs, self.config.dynamic_obstacles, 2), st.shape
            T, sp = st.shape[0], st.ctypes.data_as(C.c_void_p)
        if reset_tape is not None:
            rt = np.ascontiguousarray(np.asarray(reset_tape), dtype=np.uint32)
            width = 4 + 2 * attempts * self.config.static_obstacles + 2 * self.config.dynamic_obstacles
            assert rt.shape[1:] == (self.num_envs, width), (rt.shape, width)
            E, rp = rt.shape[0], rt.ctypes.data_as(C.c_void_p)
        check(LIB.ballenv_set_draw_tape(self._h, sp, T, rp, E, attempts))

    def stats(self) -> Dict[str, float]:
        out = (C.c_double * L.NUM_STATS)()
        check(LIB.ballenv_stats(self._h, out, self._stream()))
        return {name: out[i] for i, name in enumerate(L.STAT_NAMES)}

    @property
    def stats_tensor(self) -> torch.Tensor:
        """Device view of the statistics vector (float64 [16]); all-reduce this across GPUs."""
        return self.state_views["stats"]

    def reset_stats(self):
        check(LIB.ballenv_stats_reset(self._h, self._stream()))

    def error_flags(self) -> int:
        out = C.c_uint32(0)
        check(LIB.ballenv_error_flags(self._h, C.byref(out), self._stream()))
        return int(out.value)

    def kernel_variant(self, n_steps: int = 1, action_dtype=torch.int64) -> str:
        """Which kernel step() (n_steps = 1) / step_many() launches for index actions of this dtype:
        "lean" (thread per environment), "roles" (block of roles, production specialisation) or "generic"."""
        kind = {torch.int64: L.ACT_INDEX_I64, torch.int32: L.ACT_INDEX_I32, torch.uint8: L.ACT_INDEX_U8,
                torch.float32: L.ACT_XY_F32, torch.float64: L.ACT_XY_F64}[action_dtype]
        return ("generic", "roles", "lean", "lean")[check(LIB.ballenv_kernel_variant(self._h, kind, int(n_steps)))]

    def kernel_lanes(self, n_steps: int = 1) -> int:
        """Lanes per environment of the lean kernel step() / step_many() launches (1 or 2; 0: another kernel)."""
        v = check(LIB.ballenv_kernel_variant(self._h, L.ACT_INDEX_I64, int(n_steps)))
        return {L.KERNEL_LEAN: 1, L.KERNEL_LEAN2: 2}.get(v, 0)

    @property
    def launch_count(self) -> int:
        return int(LIB.ballenv_launch_count(self._h))
