"""Drivers for the reference's older scripts, which observe the env through the 29-float block-count vector
(``BallVecEnv.block_counts``, ``prep_state2`` of examples/ball_env_reinforce.py:130-172) - SURVEY 8f #3.

* ``BlockPolicy``      the 29 -> 128 -> 128 -> 9 net of examples/ball_env_reinforce.py:57-74 (``relu`` on every layer,
                       softmax) and, with ``logits=True``, of examples/train_supervise.py:11-30 (no ``relu`` on the
                       last layer, raw scores).  Parameter names ``affine1/2/3`` as in the reference, so the shipped
                       ``stored_models/supervised/*.pth`` and REINFORCE checkpoints of that shape load.
* ``train_supervised`` examples/train_supervise.py:68-121: cross entropy, SGD(lr 1e-2), minibatches of 200 in file order.
* ``rollout_blocks``   policy-in-the-loop stepping of N GPU environments on that observation (the loop of
                       examples/ball_env_reinforce.py:260-330 without rendering), no host synchronisation inside.
* ``reinforce_loss``   examples/ball_env_reinforce.py:88-104 ``finish_episode`` batched over environments: discounted
                       returns per environment (reset at episode ends), normalised over the batch, -log_prob * return.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .a2c import discounted_returns
from .vec_env import BallVecEnv


class BlockPolicy(nn.Module):
    def __init__(self, logits: bool = False):
        super().__init__()
        self.affine1 = nn.Linear(29, 128)
        self.affine2 = nn.Linear(128, 128)
        self.affine3 = nn.Linear(128, 9)
        self.logits = logits

    def forward(self, x):
        x = F.relu(self.affine1(x))
        x = F.relu(self.affine2(x))
        x = self.affine3(x)
        if self.logits:
            return x                                   # train_supervise.py:24-28
        return F.softmax(F.relu(x), dim=-1)            # ball_env_reinforce.py:69-74


def train_supervised(model: BlockPolicy, x: torch.Tensor, y: torch.Tensor, epochs: int, batch_size: int = 200,
                     lr: float = 1e-2) -> float:
    """-> mean loss of the last epoch.  ``model`` must return scores (``logits=True``)."""
    loss_fn = nn.CrossEntropyLoss()
    opt = torch.optim.SGD(model.parameters(), lr=lr)
    last = float("nan")
    for _ in range(epochs):
        total, batches = 0.0, 0
        for i in range(0, len(x), batch_size):
            loss = loss_fn(model(x[i:i + batch_size]), y[i:i + batch_size])
            opt.zero_grad()
            loss.backward()
            opt.step()
            total += float(loss.detach())
            batches += 1
        last = total / max(batches, 1)
    return last


def rollout_blocks(env: BallVecEnv, policy: BlockPolicy, n_steps: int, greedy: bool = False,
                   generator: Optional[torch.Generator] = None) -> Dict[str, torch.Tensor]:
    """-> dict(obs [N, 29] after the last step, log_prob [T, N], reward [T, N], done [T, N] bool, action [T, N])."""
    log_probs, rewards, dones, actions = [], [], [], []
    obs = env.block_counts()
    for _ in range(n_steps):
        out = policy(obs)
        probs = F.softmax(out, dim=-1) if policy.logits else out
        if greedy:
            action = probs.argmax(dim=-1)
        else:
            action = torch.multinomial(probs + 1e-30, 1, generator=generator).squeeze(-1)
        log_probs.append(torch.log(probs.gather(-1, action.unsqueeze(-1)).squeeze(-1) + 1e-30))
        _, reward, done, _ = env.step(action)
        rewards.append(reward.clone())
        dones.append(done.clone())
        actions.append(action)
        obs = env.block_counts()
    return dict(obs=obs, log_prob=torch.stack(log_probs), reward=torch.stack(rewards), done=torch.stack(dones),
                action=torch.stack(actions))


def reinforce_loss(batch: Dict[str, torch.Tensor], gamma: float = 0.99) -> torch.Tensor:
    ret = discounted_returns(batch["reward"], batch["done"], gamma)
    ret = (ret - ret.mean()) / (ret.std() + torch.finfo(torch.float32).eps)
    return -(batch["log_prob"] * ret).sum() / batch["log_prob"].shape[1]
