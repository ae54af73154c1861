"""Batched actor-critic rollout driver for the vector environment (BASELINE.json config 5).

The reference's training loop (examples/ball_cnn_ac3.py:528-646) drives ONE environment: per step it builds the
observation on the CPU (prep_state4, :560), copies it to the device (:412), runs ``Policy(window)`` (:109-146),
samples a 9-way ``Categorical`` and pulls the action back with ``.item()`` (:210-220, a device->host sync per
step), then calls ``env.step(move_list[action])`` (:588).  Here the same loop runs for N environments with nothing
leaving the GPU: the fused step kernel returns the observation tensor the policy consumes in place, the sampled
int64 indices are the kernel's action input, rewards / dones are accumulated on the device and finished
environments are reset inside the step launch.

``Policy`` keeps the reference's architecture and parameter names (``fc1``, ``action_head``, ``value_head``), so the
state_dicts the reference saves (:637-639; e.g. ``fc1`` 128x29 for WINDOW=5, 208x104 for WINDOW=10) load unchanged.
``a2c_loss`` is the batched form of ``finish_episode`` (:222-246).  This module is plumbing around the hot path
(plain PyTorch); the hot path itself is the CUDA kernel behind ``BallVecEnv.step``.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .vec_env import BallVecEnv

EPS = float(torch.finfo(torch.float32).eps)      # np.finfo(np.float32).eps.item(), examples/ball_cnn_ac3.py:76


class Policy(nn.Module):
    """examples/ball_cnn_ac3.py:109-146: (4 + W*W) -> hidden (128 for W=5, 208 for W=10) -> {9 action logits, 1 value}."""

    def __init__(self, window: int, hidden: Optional[int] = None):
        super().__init__()
        n_in = 4 + window * window
        if hidden is None:
            hidden = {5: 128, 10: 208}.get(window, 2 * n_in)      # 2 * inputs is the reference's commented default (:115)
        self.hidden_layer = hidden
        self.fc1 = nn.Linear(n_in, hidden)
        self.action_head = nn.Linear(hidden, 9)
        self.value_head = nn.Linear(hidden, 1)

    def forward(self, x):
        x = F.relu(self.fc1(x))
        return F.softmax(self.action_head(x), dim=-1), self.value_head(x)


class PolicyCNN(nn.Module):
    """examples/ball_cnn_ac3.py:80-106 (the same class in ball_cnn_reinforce.py): three 5 x 5 stride-2 convolutions with
    batch norm over the 3 x 40 x 40 patch (``BallVecEnv.rgb_patches``), the 128 features joined with the 4 goal-quadrant
    floats (``obs[:, :4]``), then the 9-way action head and the value head.  Parameter names as in the reference."""

    def __init__(self):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 16, kernel_size=5, stride=2)
        self.bn1 = nn.BatchNorm2d(16)
        self.conv2 = nn.Conv2d(16, 32, kernel_size=5, stride=2)
        self.bn2 = nn.BatchNorm2d(32)
        self.conv3 = nn.Conv2d(32, 32, kernel_size=5, stride=2)
        self.bn3 = nn.BatchNorm2d(32)
        self.action_head = nn.Linear(132, 9)
        self.value_head = nn.Linear(132, 1)

    def forward(self, x, y):
        x = F.relu(self.bn1(self.conv1(x)))
        x = F.relu(self.bn2(self.conv2(x)))
        x = F.relu(self.bn3(self.conv3(x)))
        x = torch.cat((x.view(x.size(0), -1), y), 1)
        return F.softmax(self.action_head(x), dim=-1), self.value_head(x)


def rollout(env: BallVecEnv, policy: Policy, n_steps: int, obs: Optional[torch.Tensor] = None, greedy: bool = False,
            generator: Optional[torch.Generator] = None) -> Dict[str, torch.Tensor]:
    """n_steps of policy-in-the-loop stepping for all environments; no host synchronisation inside the loop.

    -> dict(obs [N, row] (the observation after the last step), log_prob [T, N], value [T, N], reward [T, N],
            done [T, N] bool, action [T, N] int64)."""
    if obs is None:
        obs = env.observe()
    log_probs, values, rewards, dones, actions = [], [], [], [], []
    for _ in range(n_steps):
        # The observation is a view of an env-owned double buffer that the kernel rewrites two steps later behind
        # autograd's back (a raw launch does not bump the tensor's version counter): give the graph its own copy, or
        # fc1's weight gradient would be computed from later observations.
        x = obs.float()
        if x.data_ptr() == obs.data_ptr():
            x = x.clone()
        probs, value = policy(x)
        if greedy:
            action = probs.argmax(dim=-1)
        else:
            action = torch.multinomial(probs, 1, generator=generator).squeeze(-1)      # Categorical(probs).sample()
        log_probs.append(torch.log(probs.gather(-1, action.unsqueeze(-1)).squeeze(-1)))
        values.append(value.squeeze(-1))
        obs, reward, done, _ = env.step(action)      # views of env-owned buffers: copy what is kept
        rewards.append(reward.clone())
        dones.append(done.clone())
        actions.append(action)
    return dict(obs=obs, log_prob=torch.stack(log_probs), value=torch.stack(values), reward=torch.stack(rewards),
                done=torch.stack(dones), action=torch.stack(actions))


class GraphedRollout:
    """The whole T-step policy-in-the-loop rollout as ONE CUDA graph.

    The per-step work of the reference loop (examples/ball_cnn_ac3.py:553-613) is a dozen tiny kernels (the MLP,
    softmax, sampling, the fused environment step); launched one by one they are launch-bound.  Here the T steps -
    policy forward, ``multinomial`` sampling and ``ballenv_step`` writing into fixed buffers - are captured once and
    replayed, so the GPU runs them back to back.  The rollout is recorded without autograd; ``evaluate`` recomputes
    log-probabilities and values of the stored (observation, action) pairs in one batched forward pass for the
    update (same numbers as the step-by-step values, since the weights do not change inside a rollout).
    """

    def __init__(self, env: BallVecEnv, policy: Policy, n_steps: int, greedy: bool = False):
        self.env, self.policy, self.n_steps, self.greedy = env, policy, n_steps, greedy
        n, row, dev = env.num_envs, env.obs_row, env.device
        self.obs = torch.zeros((n_steps + 1, n, row), dtype=torch.float32, device=dev)   # obs[t] is what step t sees
        self.action = torch.zeros((n_steps, n), dtype=torch.int64, device=dev)
        self.reward = torch.zeros((n_steps, n), dtype=torch.float32, device=dev)
        self.done = torch.zeros((n_steps, n), dtype=torch.uint8, device=dev)
        self.graph = None

    def _body(self):
        for t in range(self.n_steps):
            probs, _ = self.policy(self.obs[t])
            a = probs.argmax(dim=-1) if self.greedy else torch.multinomial(probs, 1).squeeze(-1)
            self.action[t].copy_(a)
            self.env.step_into(self.action[t], self.obs[t + 1], self.reward[t], self.done[t])

    @torch.no_grad()
    def run(self, first_obs: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        """One rollout.  ``first_obs`` defaults to the last observation of the previous rollout (or the current
        observation of the environment on the first call)."""
        if first_obs is None:
            first_obs = self.obs[self.n_steps].clone() if self.graph is not None else self.env.observe()
        self.obs[0].copy_(first_obs)
        if self.graph is None:
            side = torch.cuda.Stream(device=self.env.device)
            side.wait_stream(torch.cuda.current_stream(self.env.device))
            with torch.cuda.stream(side):          # warm-up on a side stream, as graph capture requires
                self._body()
            torch.cuda.current_stream(self.env.device).wait_stream(side)
            self.obs[0].copy_(first_obs)           # the warm-up advanced the environments for real: keep that rollout
            out_first = dict(obs=self.obs.clone(), action=self.action.clone(), reward=self.reward.clone(),
                             done=self.done.clone().bool())
            self.graph = torch.cuda.CUDAGraph()
            self.obs[0].copy_(self.obs[self.n_steps])
            with torch.cuda.graph(self.graph):
                self._body()
            # the capture itself does not execute anything; hand back the warm-up rollout this time
            return out_first
        self.graph.replay()
        return dict(obs=self.obs, action=self.action, reward=self.reward, done=self.done.bool())

    def evaluate(self, batch: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        """log-probabilities and values of the stored pairs, with autograd, in one batched forward pass."""
        T, n = batch["action"].shape
        probs, value = self.policy(batch["obs"][:T].reshape(T * n, -1))
        logp = torch.log(probs.gather(-1, batch["action"].reshape(-1, 1)).squeeze(-1)).view(T, n)
        return dict(log_prob=logp, value=value.view(T, n), reward=batch["reward"], done=batch["done"])


class FusedRollout(GraphedRollout):
    """The T-step policy-in-the-loop rollout as ONE KERNEL LAUNCH (``BallVecEnv.rollout_policy``): the environments'
    own lanes evaluate the MLP and draw the action between two steps, so nothing but observations, actions, rewards
    and dones touches device memory inside the loop.  Same buffers and ``evaluate`` as GraphedRollout; actions are
    drawn from the env's Philox stream (inverse CDF of the same softmax) instead of ``torch.multinomial``.
    ``keep_policy_out``: also keep the action probabilities and the value of every step's forward pass
    (``policy_out`` [T, N, 10]) - with unchanged weights exactly what an update would recompute (FusedUpdate takes it)."""

    def __init__(self, env: BallVecEnv, policy: Policy, n_steps: int, greedy: bool = False, keep_policy_out: bool = False):
        super().__init__(env, policy, n_steps, greedy)
        self.policy_out = (torch.zeros((n_steps, env.num_envs, 10), dtype=torch.float32, device=env.device)
                           if keep_policy_out else None)

    @torch.no_grad()
    def run(self, first_obs: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        if first_obs is None:
            first_obs = self.obs[self.n_steps].clone() if self.graph is not None else self.env.observe()
        self.graph = True      # (no graph to capture: marks that obs[n_steps] holds the last observation)
        self.obs[0].copy_(first_obs)
        self.env.rollout_policy(self.policy, self.n_steps, self.obs[0], self.obs[1:], self.action, self.reward, self.done,
                                greedy=self.greedy, policy_out=self.policy_out)
        out = dict(obs=self.obs, action=self.action, reward=self.reward, done=self.done.bool())
        if self.policy_out is not None:
            out["policy_out"] = self.policy_out
        return out


def train_fused(env: BallVecEnv, policy: Policy, iterations: int, n_steps: int = 32, gamma: float = 0.99,
                lr: float = 1e-3, log=None):
    """train_graphed with the rollout in one launch (FusedRollout)."""
    return train_graphed(env, policy, iterations, n_steps, gamma, lr, log, rollout_cls=FusedRollout)


class FusedUpdate:
    """``a2c_loss(...).backward()`` for ``Policy`` without autograd: the loss and the gradients of all six parameter
    tensors from hand-written kernels (ballenv_a2c_grads - forward and backward per sample on chip, per-unit sums
    without atomics, a deterministic second pass), written straight into ``param.grad``.  The returns are prepared as
    ``a2c_loss`` does (discounted, restarted at episode ends, bootstrapped, normalised over the batch); the advantage
    is a constant, as in the reference (``value.item()``, examples/ball_cnn_ac3.py:234)."""

    def __init__(self, policy: Policy, n_samples: int):
        import ctypes as C
        from . import _lib as L
        self.policy, self.n_samples = policy, int(n_samples)
        ps = [policy.fc1.weight, policy.fc1.bias, policy.action_head.weight, policy.action_head.bias,
              policy.value_head.weight, policy.value_head.bias]
        for q in ps:
            if q.dtype != torch.float32 or not q.is_cuda or not q.is_contiguous():
                raise ValueError("FusedUpdate needs contiguous float32 CUDA parameters")
            if q.grad is None:
                q.grad = torch.zeros_like(q)          # fixed buffers: the kernels overwrite them every call
        self.params = ps
        self.device = ps[0].device
        hidden, n_in = policy.fc1.weight.shape
        with torch.cuda.device(self.device):
            nbytes = L.check(L.LIB.ballenv_a2c_workspace_bytes(n_in, hidden, self.n_samples))
        self.workspace = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        self.loss = torch.zeros(1, dtype=torch.float32, device=self.device)
        self.stats = torch.zeros(2, dtype=torch.float32, device=self.device)     # {mean, std + eps} of raw returns
        self._u = L.BallenvA2CUpdate(n_in, hidden, *[q.data_ptr() for q in ps], *[q.grad.data_ptr() for q in ps],
                                     self.loss.data_ptr(), None, None)
        self._C, self._L = C, L

    def grads(self, obs: torch.Tensor, action: torch.Tensor, returns: torch.Tensor, normalise: bool = False,
              policy_out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """obs [S, row] float32, action [S] int64, returns [S] float32 -> loss (device scalar view); the parameters'
        ``.grad`` hold d loss / d parameter afterwards.  ``returns`` are the normalised returns, or with
        ``normalise=True`` the raw discounted ones: their mean and unbiased std + eps are taken here (one reduction) and
        the kernel forms (R - mean) / (std + eps) itself (examples/ball_cnn_ac3.py:231-232).  ``policy_out`` [S, 10]
        (FusedRollout's, taken with the SAME weights): the probabilities and values are read instead of recomputed."""
        C, L = self._C, self._L
        if policy_out is not None:
            if not (policy_out.is_contiguous() and policy_out.dtype == torch.float32 and policy_out.numel() == self.n_samples * 10):
                raise ValueError("policy_out must be contiguous float32 [S, 10]")
            self._u.policy_out = policy_out.data_ptr()
        else:
            self._u.policy_out = None
        if normalise:
            var, mean = torch.var_mean(returns.reshape(-1))
            self.stats[0] = mean
            self.stats[1] = var.sqrt() + EPS
            self._u.returns_stats = self.stats.data_ptr()
        else:
            self._u.returns_stats = None
        S = self.n_samples
        if not (obs.is_contiguous() and action.is_contiguous() and returns.is_contiguous() and obs.numel() == S * obs.shape[-1]
                and action.numel() == S and returns.numel() == S and obs.dtype == torch.float32
                and action.dtype == torch.int64 and returns.dtype == torch.float32):
            raise ValueError("grads needs contiguous obs [S, row] f32, action [S] i64, returns [S] f32")
        for q, g in zip(self.params, (q.grad for q in self.params)):   # (an optimizer's zero_grad(set_to_none=True) drops them)
            if g is None:
                raise RuntimeError("keep the .grad buffers: call zero_grad(set_to_none=False) or none at all")
        with torch.cuda.device(self.device):
            L.check(L.LIB.ballenv_a2c_grads(C.byref(self._u), C.c_void_p(obs.data_ptr()), C.c_void_p(action.data_ptr()),
                                            C.c_void_p(returns.data_ptr()), S, C.c_void_p(self.workspace.data_ptr()),
                                            self.workspace.numel(),
                                            C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))
        return self.loss[0]


def normalised_returns(reward: torch.Tensor, done: torch.Tensor, gamma: float, bootstrap: Optional[torch.Tensor] = None):
    """The returns ``a2c_loss`` regresses on: discounted (restarted at episode ends, bootstrapped), then normalised by
    the batch's mean and unbiased standard deviation + eps (examples/ball_cnn_ac3.py:228-232)."""
    returns = discounted_returns(reward.float(), done, gamma, bootstrap)
    return (returns - returns.mean()) / (returns.std() + EPS)


class GraphedTrainer:
    """One whole iteration of the actor-critic loop - the fused T-step rollout (one launch), the batched
    ``finish_episode`` update and the Adam step - as ONE CUDA graph, replayed per iteration: nothing is launched from
    Python inside the loop.  ``fused_update`` (default): loss and gradients come from ``FusedUpdate`` (hand-written
    forward + backward, two launches); otherwise from autograd over the stored pairs (``evaluate`` + ``a2c_loss`` +
    ``backward``: a few dozen small kernels that move the hidden activations through HBM).  Needs a configuration with
    a policy-in-the-loop kernel (``BallVecEnv.rollout_policy``)."""

    def __init__(self, env: BallVecEnv, policy: Policy, n_steps: int = 32, gamma: float = 0.99, lr: float = 1e-3,
                 fused_update: bool = True):
        self.env, self.policy, self.n_steps, self.gamma = env, policy, n_steps, gamma
        # (fused: the whole Adam step of the six tensors is one kernel)
        self.opt = torch.optim.Adam(policy.parameters(), lr=lr, capturable=True, fused=bool(fused_update))
        self.roll = FusedRollout(env, policy, n_steps, keep_policy_out=bool(fused_update))
        self.loss = torch.zeros((), dtype=torch.float32, device=env.device)
        self.update = FusedUpdate(policy, n_steps * env.num_envs) if fused_update else None
        self.graph = None

    def _body(self):
        raw = self.roll.run()
        T = self.n_steps
        if self.update is not None:
            # the update without autograd: returns as a2c_loss prepares them, then loss + gradients in two launches
            with torch.no_grad():
                _, v_last = self.policy(raw["obs"][T])
                returns = discounted_returns(raw["reward"], self.roll.done, self.gamma, bootstrap=v_last.squeeze(-1))
                loss = self.update.grads(raw["obs"][:T], raw["action"], returns, normalise=True,
                                         policy_out=raw["policy_out"])
            self.opt.step()
            self.loss.copy_(loss)
            return
        batch = self.roll.evaluate(raw)
        with torch.no_grad():
            _, v_last = self.policy(raw["obs"][T])
        loss = a2c_loss(batch, self.gamma, bootstrap=v_last.squeeze(-1))
        self.opt.zero_grad(set_to_none=True)
        loss.backward()
        self.opt.step()
        self.loss.copy_(loss.detach())

    def step(self) -> torch.Tensor:
        """One iteration (the first call runs three, eagerly, to warm up before the capture); returns the loss, a device
        scalar that the next call overwrites."""
        if self.graph is None:
            dev = self.env.device
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):       # off the capturing stream: allocations, cuBLAS handles, Adam state
                for _ in range(3):
                    self._body()
            torch.cuda.current_stream(dev).wait_stream(side)
            self.graph = torch.cuda.CUDAGraph()
            if self.update is None:
                self.opt.zero_grad(set_to_none=True)
            with torch.cuda.graph(self.graph):
                self._body()
            return self.loss
        self.graph.replay()
        return self.loss


def train_graphed(env: BallVecEnv, policy: Policy, iterations: int, n_steps: int = 32, gamma: float = 0.99,
                  lr: float = 1e-3, log=None, rollout_cls=None):
    """train() with the rollout replayed as a CUDA graph and the update computed from one batched forward pass."""
    opt = torch.optim.Adam(policy.parameters(), lr=lr)
    env.reset()
    roll = (rollout_cls or GraphedRollout)(env, policy, n_steps)
    for it in range(iterations):
        raw = roll.run()
        batch = roll.evaluate(raw)
        with torch.no_grad():
            _, v_last = policy(raw["obs"][n_steps])
        loss = a2c_loss(batch, gamma, bootstrap=v_last.squeeze(-1))
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        if log is not None:
            log(it, loss, batch)
    return policy


def discounted_returns(reward: torch.Tensor, done: torch.Tensor, gamma: float, bootstrap: Optional[torch.Tensor] = None):
    """R_t = r_t + gamma * R_{t+1}, restarted where an episode ended (examples/ball_cnn_ac3.py:228-230, per env)."""
    T = reward.shape[0]
    out = torch.empty_like(reward)
    if (reward.is_cuda and reward.dtype == torch.float32 and reward.dim() == 2 and reward.is_contiguous()
            and done.dtype in (torch.bool, torch.uint8) and done.is_contiguous() and done.shape == reward.shape
            and not reward.requires_grad and (bootstrap is None or not bootstrap.requires_grad)):
        # one launch, one thread per environment (ballenv_discounted_returns): same float32 roundings as the loop below
        import ctypes as C
        from ._lib import LIB, check
        boot = None
        if bootstrap is not None:
            boot = bootstrap.to(torch.float32).contiguous()
        with torch.cuda.device(reward.device):
            check(LIB.ballenv_discounted_returns(
                C.c_void_p(reward.data_ptr()), C.c_void_p(done.data_ptr()), None if boot is None else C.c_void_p(boot.data_ptr()),
                C.c_float(gamma), T, reward.shape[1], C.c_void_p(out.data_ptr()),
                C.c_void_p(torch.cuda.current_stream(reward.device).cuda_stream)))
        return out
    R = torch.zeros_like(reward[0]) if bootstrap is None else bootstrap
    for t in range(T - 1, -1, -1):
        R = reward[t] + gamma * R * (~done[t]).to(reward.dtype)
        out[t] = R
    return out


def finish_episode_loss(log_prob: torch.Tensor, value: torch.Tensor, rewards: torch.Tensor, gamma: float = 0.99):
    """The loss of the reference's ``finish_episode`` (examples/ball_cnn_ac3.py:222-246) for ONE finished episode:
    discounted returns from the end (:228-230), normalised by their mean and unbiased standard deviation + eps (:231-232),
    sum of -log_prob * (R - V.item()) and of smooth_l1(V, R) (:233-241).  ``log_prob``, ``value``, ``rewards``: [T].
    tests/golden/a2c_kat.npz holds what the reference's own function computes (loss and gradients).  ``a2c_loss`` below
    is the same formula over a [T, N] batch; for N = 1, no episode end inside and no bootstrap it is this function."""
    batch = dict(log_prob=log_prob.view(-1, 1), value=value.view(-1, 1), reward=rewards.view(-1, 1),
                 done=torch.zeros(rewards.numel(), 1, dtype=torch.bool, device=rewards.device))
    return a2c_loss(batch, gamma)


def a2c_loss(batch: Dict[str, torch.Tensor], gamma: float = 0.99, bootstrap: Optional[torch.Tensor] = None):
    """Batched finish_episode (examples/ball_cnn_ac3.py:222-246): returns normalised over the batch,
    policy loss -log_prob * (R - V.detach()), value loss smooth_l1(V, R), summed.

    Intended differences from the reference, which updates once per finished episode of its single environment:
    the batch is a fixed-length slice of N environments' trajectories, so (a) returns restart wherever an episode
    ended inside the slice (``done``) and may be bootstrapped with V(s_T) where it did not, (b) mean / std are taken
    over the whole [T, N] batch instead of one episode.  With N = 1, one whole episode and no bootstrap both
    differences vanish (``finish_episode_loss``; pinned by tests/test_a2c_driver.py against the reference's function)."""
    returns = discounted_returns(batch["reward"].float(), batch["done"], gamma, bootstrap)
    returns = (returns - returns.mean()) / (returns.std() + EPS)
    advantage = returns - batch["value"].detach()
    policy_loss = (-batch["log_prob"] * advantage).sum()
    value_loss = F.smooth_l1_loss(batch["value"], returns, reduction="sum")
    return policy_loss + value_loss


def train(env: BallVecEnv, policy: Policy, iterations: int, n_steps: int = 32, gamma: float = 0.99, lr: float = 1e-3,
          generator: Optional[torch.Generator] = None, log=None):
    """Adam(lr=1e-3) as examples/ball_cnn_ac3.py:505; one update per n_steps-step rollout of all environments."""
    opt = torch.optim.Adam(policy.parameters(), lr=lr)
    obs = env.reset()
    for it in range(iterations):
        batch = rollout(env, policy, n_steps, obs=obs, generator=generator)
        obs = batch["obs"]
        with torch.no_grad():
            _, v_last = policy(obs.float())
        loss = a2c_loss(batch, gamma, bootstrap=v_last.squeeze(-1))
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        if log is not None:
            log(it, loss, batch)
    return policy
