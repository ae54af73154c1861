"""Config 5: the actor-critic rollout loop of examples/ball_cnn_ac3.py driving 16K GPU environments end to end."""
import numpy as np
import pytest
import torch


def test_policy_has_the_reference_parameter_layout():
    """state_dicts saved by the reference (examples/ball_cnn_ac3.py:637-639) must load: fc1 128x29 / 208x104."""
    from gym_ballenv_b200.a2c import Policy
    for w, hidden in ((5, 128), (10, 208)):
        sd = Policy(w).state_dict()
        assert list(sd) == ["fc1.weight", "fc1.bias", "action_head.weight", "action_head.bias",
                            "value_head.weight", "value_head.bias"]
        assert tuple(sd["fc1.weight"].shape) == (hidden, 4 + w * w) and tuple(sd["action_head.weight"].shape) == (9, hidden)
    probs, value = Policy(5)(torch.zeros(3, 29))
    assert probs.shape == (3, 9) and value.shape == (3, 1) and torch.allclose(probs.sum(-1), torch.ones(3))


def test_discounted_returns_restart_at_episode_ends():
    from gym_ballenv_b200.a2c import a2c_loss, discounted_returns
    r = torch.tensor([[1.0], [2.0], [3.0], [4.0]])
    d = torch.tensor([[False], [True], [False], [False]])
    out = discounted_returns(r, d, 0.5)
    assert out.flatten().tolist() == [1.0 + 0.5 * 2.0, 2.0, 3.0 + 0.5 * 4.0, 4.0]        # finish_episode :228-230 per episode
    batch = dict(reward=r, done=d, value=torch.zeros(4, 1, requires_grad=True), log_prob=torch.zeros(4, 1, requires_grad=True))
    assert torch.isfinite(a2c_loss(batch, 0.5))


@pytest.mark.gpu
def test_closed_loop_16k_envs_against_the_c_oracle():
    """16384 environments, the policy in the loop (greedy actions so both sides see the same indices): every
    observation the policy consumes is bit-identical to the C oracle's, rewards to 1e-5."""
    from gym_ballenv_b200 import BallVecEnv
    from gym_ballenv_b200.a2c import Policy
    from oracle.ballenv_oracle import OracleConfig
    from oracle.c_oracle import COracleVec
    n, T, seed = 16384, 40, 31
    torch.manual_seed(0)
    policy = Policy(5).cuda()
    env = BallVecEnv(n, window=5, seed=seed, max_episode_steps=15)
    c = COracleVec(OracleConfig(window=5, max_episode_steps=15), seed, n)
    obs = env.reset()
    c.reset()
    with torch.no_grad():
        for t in range(T):
            assert np.array_equal(obs.cpu().numpy(), c.observe()), t
            probs, _ = policy(obs.float())
            action = probs.argmax(-1)
            obs, reward, done, _ = env.step(action)
            r, d, f = c.step(action.cpu().numpy())
            assert np.array_equal(done.cpu().numpy(), d), t
            np.testing.assert_allclose(reward.cpu().numpy(), r, rtol=1e-5, atol=0)
    assert env.error_flags() == 0


@pytest.mark.gpu
def test_train_loop_runs_and_updates_the_policy():
    from gym_ballenv_b200 import BallVecEnv
    from gym_ballenv_b200.a2c import Policy, train
    torch.manual_seed(1)
    env = BallVecEnv(16384, window=5, seed=3)
    policy = Policy(5).cuda()
    before = policy.fc1.weight.detach().clone()
    losses = []
    train(env, policy, iterations=3, n_steps=16, generator=torch.Generator(device="cuda").manual_seed(2),
          log=lambda it, loss, batch: losses.append(float(loss.detach())))
    assert len(losses) == 3 and all(np.isfinite(losses))
    assert not torch.equal(before, policy.fc1.weight.detach())
    assert env.stats()["steps"] == 16384 * 16 * 3 and env.error_flags() == 0


@pytest.mark.gpu
def test_graphed_rollout_equals_the_eager_loop():
    """The T-step rollout replayed as one CUDA graph (greedy actions, so both sides are deterministic) visits the same
    observations, actions, rewards and dones as the step-by-step loop, over several replays with resets."""
    from gym_ballenv_b200 import BallVecEnv
    from gym_ballenv_b200.a2c import GraphedRollout, Policy, rollout
    torch.manual_seed(3)
    policy = Policy(5).cuda()
    n, T = 4096, 16
    e1 = BallVecEnv(n, window=5, seed=9, max_episode_steps=11)
    e2 = BallVecEnv(n, window=5, seed=9, max_episode_steps=11)
    e1.reset()
    obs2 = e2.reset().clone()
    g = GraphedRollout(e1, policy, T, greedy=True)
    with torch.no_grad():
        for it in range(4):
            raw = g.run()
            ref = rollout(e2, policy, T, obs=obs2, greedy=True)
            obs2 = ref["obs"].clone()
            assert torch.equal(raw["action"], ref["action"]), it
            assert torch.equal(raw["reward"], ref["reward"]) and torch.equal(raw["done"], ref["done"]), it
            assert torch.equal(raw["obs"][T], ref["obs"]), it
    ev = g.evaluate(raw)
    assert ev["log_prob"].requires_grad and ev["value"].shape == (T, n)
    assert e1.stats() == e2.stats() or abs(e1.stats()["return_sum"] - e2.stats()["return_sum"]) < 1e-6 * abs(e2.stats()["return_sum"])


@pytest.mark.gpu
def test_graphed_training_runs():
    from gym_ballenv_b200 import BallVecEnv
    from gym_ballenv_b200.a2c import Policy, train_graphed
    torch.manual_seed(1)
    env = BallVecEnv(8192, window=5, seed=3)
    policy = Policy(5).cuda()
    before = policy.fc1.weight.detach().clone()
    losses = []
    train_graphed(env, policy, iterations=4, n_steps=8, log=lambda it, loss, batch: losses.append(float(loss.detach())))
    assert len(losses) == 4 and all(np.isfinite(losses)) and not torch.equal(before, policy.fc1.weight.detach())
    assert env.stats()["steps"] == 8192 * 8 * 4 and env.error_flags() == 0
