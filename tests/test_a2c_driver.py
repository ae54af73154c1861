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


def _reference_policy():
    """Policy(5) holding the state_dict the reference ships (episode_2500.pth of its README run), from the fixture."""
    from gym_ballenv_b200.a2c import Policy
    from helpers import load_golden
    z, meta = load_golden("a2c_kat")
    policy = Policy(5)
    sd = {k: torch.from_numpy(z["ckpt_" + k.replace(".", "_")]) for k in policy.state_dict()}
    policy.load_state_dict(sd)        # same parameter names and shapes as examples/ball_cnn_ac3.py:109-146
    return policy, z, meta


def test_finish_episode_loss_and_gradients_match_the_reference():
    """tests/golden/a2c_kat.npz: the reference's own Policy + finish_episode (AST-lifted, run unedited by
    oracle/gen_golden.py) on one recorded 24-step episode with the shipped checkpoint - forward pass, loss and the
    gradient of every parameter."""
    from gym_ballenv_b200.a2c import finish_episode_loss
    policy, z, meta = _reference_policy()
    x = torch.from_numpy(z["states"])
    probs, value = policy(x)
    np.testing.assert_allclose(probs.detach().numpy(), z["probs"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(value.detach().numpy()[:, 0], z["values"], rtol=1e-5, atol=1e-7)
    actions = torch.from_numpy(z["actions"])
    assert torch.equal(probs.argmax(-1), actions)                       # the recorded episode was greedy
    logp = torch.log(probs.gather(-1, actions.view(-1, 1)).squeeze(-1))
    loss = finish_episode_loss(logp, value.squeeze(-1), torch.from_numpy(z["rewards"]).float(), meta["gamma"])
    assert float(loss.detach()) == pytest.approx(float(z["loss"]), rel=2e-5)
    loss.backward()
    for name, prm in policy.named_parameters():
        ref = z["grad_" + name.replace(".", "_")]
        np.testing.assert_allclose(prm.grad.numpy(), ref, rtol=2e-3, atol=2e-5 * float(np.abs(ref).max()))


def test_shipped_checkpoint_drives_the_oracle_like_the_reference():
    """tests/golden/rollout_checkpoint.npz: the reference environment driven in closed loop by the reference's own
    network (shipped checkpoint, greedy) through its own prep_state4.  Here: the CPU oracle port observed through its
    window raster, this package's Policy with the same weights - same actions at every step of every episode."""
    from helpers import load_golden, oracle_config
    from oracle import draws as D
    from oracle.ballenv_oracle import OracleVec
    policy, _, _ = _reference_policy()
    z, meta = load_golden("rollout_checkpoint")
    n, T = meta["n_envs"], meta["T"]
    vec = OracleVec(oracle_config(meta["cfg"], 5, meta["max_episode_steps"]), D.PhiloxDraws(meta["seed"]), n, meta["g0"])
    vec.reset()
    with torch.no_grad():
        for t in range(T):
            obs = torch.tensor(vec.observe(), dtype=torch.float32)
            a = policy(obs)[0].argmax(-1)
            assert a.tolist() == list(z["rec_actions"][t]), t
            rew, done, flags = vec.step(a.tolist())
            assert rew == list(z["rec_reward"][t]) and [int(d) for d in done] == list(z["rec_done"][t]), t
    assert vec.stats["episodes"] == meta["stats"]["episodes"]


@pytest.mark.gpu
def test_shipped_checkpoint_closed_loop_on_the_gpu():
    """The same closed loop on the CUDA path: BallVecEnv observations -> Policy (shipped weights, on the device) ->
    greedy action -> ballenv_step, against the reference's recorded episodes (actions, rewards, dones, quadrant bits and
    window rows), auto-resets included."""
    from gym_ballenv_b200 import BallVecEnv, EnvConfig
    from helpers import load_golden
    policy, _, _ = _reference_policy()
    policy = policy.cuda()
    z, meta = load_golden("rollout_checkpoint")
    n, T = meta["n_envs"], meta["T"]
    env = BallVecEnv(n, window=5, config=EnvConfig(), seed=meta["seed"], max_episode_steps=meta["max_episode_steps"],
                     global_env_offset=meta["g0"])
    obs = env.reset()
    with torch.no_grad():
        for t in range(T):
            a = policy(obs.float())[0].argmax(-1)
            assert a.cpu().tolist() == list(z["rec_actions"][t]), t
            obs, rew, done, _ = env.step(a)
            np.testing.assert_allclose(rew.cpu().numpy(), z["rec_reward"][t], rtol=1e-5, atol=0)
            assert done.cpu().numpy().astype(np.uint8).tolist() == list(z["rec_done"][t]), t
            o = obs.cpu().numpy()
            assert np.array_equal(o[:, :4].argmax(1), z["rec_quadrant"][t]), t
            rows = (o[:, 4:].reshape(n, 5, 5) * (1 << np.arange(5))).sum(-1).astype(np.uint32)
            assert np.array_equal(rows, z["rec_rows5"][t]), t
    assert env.stats()["episodes"] == meta["stats"]["episodes"] and env.error_flags() == 0
    env.close()


@pytest.mark.gpu
@pytest.mark.parametrize("lanes", [1, 2])
def test_shipped_checkpoint_in_the_fused_rollout(lanes, monkeypatch):
    """The same recorded episodes against ONE launch: the shipped checkpoint evaluated inside the rollout kernel
    (ballenv_rollout_policy, greedy) - the actions, rewards, dones and observations the reference's own network and
    environment produced, step by step."""
    from gym_ballenv_b200 import BallVecEnv, EnvConfig
    from gym_ballenv_b200.a2c import FusedRollout
    from helpers import load_golden
    policy, _, _ = _reference_policy()
    policy = policy.cuda()
    z, meta = load_golden("rollout_checkpoint")
    n, T = meta["n_envs"], meta["T"]
    monkeypatch.setenv("BALLENV_LEAN_G", str(lanes))
    env = BallVecEnv(n, window=5, config=EnvConfig(), seed=meta["seed"], max_episode_steps=meta["max_episode_steps"],
                     global_env_offset=meta["g0"])
    monkeypatch.delenv("BALLENV_LEAN_G")
    env.reset()
    raw = FusedRollout(env, policy, T, greedy=True, keep_policy_out=(lanes == 2)).run()
    assert np.array_equal(raw["action"].cpu().numpy(), z["rec_actions"][:T])
    if lanes == 2:      # the probabilities and values it kept are the network's on the observations it acted on
        with torch.no_grad():
            probs, value = policy(raw["obs"][:T].reshape(T * n, -1))
        po = raw["policy_out"].reshape(T * n, 10)
        assert torch.allclose(po[:, :9], probs, rtol=1e-4, atol=1e-6) and torch.allclose(po[:, 9], value[:, 0], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(raw["reward"].cpu().numpy(), z["rec_reward"][:T], rtol=1e-5, atol=0)
    assert np.array_equal(raw["done"].cpu().numpy().astype(np.uint8), np.asarray(z["rec_done"][:T], dtype=np.uint8))
    o = raw["obs"][1:].cpu().numpy()
    assert np.array_equal(o[:, :, :4].argmax(2), z["rec_quadrant"][:T])
    rows = (o[:, :, 4:].reshape(T, n, 5, 5) * (1 << np.arange(5))).sum(-1).astype(np.uint32)
    assert np.array_equal(rows, z["rec_rows5"][:T])
    assert env.stats()["episodes"] == meta["stats"]["episodes"] and env.error_flags() == 0
    env.close()


@pytest.mark.gpu
def test_eager_rollout_gradients_do_not_alias_the_env_buffers():
    """rollout() feeds the policy env-owned observation buffers that the kernel rewrites two steps later; the autograd
    graph must hold its own copies.  The gradients of a rollout's loss equal those recomputed from cloned
    (observation, action) pairs in one batched forward pass."""
    from gym_ballenv_b200 import BallVecEnv
    from gym_ballenv_b200.a2c import Policy, a2c_loss, rollout
    torch.manual_seed(4)
    n, T = 2048, 12
    env = BallVecEnv(n, window=5, seed=6)
    twin = BallVecEnv(n, window=5, seed=6)
    policy = Policy(5).cuda()
    obs = env.reset()
    obs_t = twin.reset().clone()
    batch = rollout(env, policy, T, obs=obs, greedy=True)
    loss = a2c_loss(batch, 0.99)
    policy.zero_grad()
    loss.backward()
    got = {k: p.grad.clone() for k, p in policy.named_parameters()}
    # recompute: same trajectory on a twin env (greedy actions are deterministic), observations cloned per step
    seen = []
    with torch.no_grad():
        for t in range(T):
            seen.append(obs_t.clone())
            obs_t, _, _, _ = twin.step(batch["action"][t])
            obs_t = obs_t.clone()
    probs, value = policy(torch.stack(seen).view(T * n, -1))
    logp = torch.log(probs.gather(-1, batch["action"].reshape(-1, 1)).squeeze(-1)).view(T, n)
    ref = a2c_loss(dict(log_prob=logp, value=value.view(T, n), reward=batch["reward"], done=batch["done"]), 0.99)
    policy.zero_grad()
    ref.backward()
    assert float(loss.detach()) == pytest.approx(float(ref.detach()), rel=1e-5)
    for k, p in policy.named_parameters():
        torch.testing.assert_close(got[k], p.grad, rtol=1e-3, atol=1e-4 * float(p.grad.abs().max()))
    env.close()
    twin.close()


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


@pytest.mark.gpu
def test_hand_written_update_matches_the_references_finish_episode():
    """The same fixture against ballenv_a2c_grads (a2c.FusedUpdate): the reference's own finish_episode on its recorded
    24-step episode with the shipped checkpoint - loss and the gradient of every parameter from the hand-written
    forward + backward kernels, no autograd anywhere."""
    from gym_ballenv_b200.a2c import FusedUpdate, normalised_returns
    policy, z, meta = _reference_policy()
    policy = policy.to("cuda:0")
    x = torch.from_numpy(z["states"]).to("cuda:0").float().contiguous()
    actions = torch.from_numpy(z["actions"]).to("cuda:0").long().contiguous()
    rewards = torch.from_numpy(z["rewards"]).to("cuda:0").float().view(-1, 1)
    done = torch.zeros_like(rewards, dtype=torch.bool)
    returns = normalised_returns(rewards, done, meta["gamma"]).view(-1).contiguous()      # :228-232, one episode
    upd = FusedUpdate(policy, x.shape[0])
    loss = upd.grads(x, actions, returns)
    assert float(loss) == pytest.approx(float(z["loss"]), rel=2e-5)
    for name, prm in policy.named_parameters():
        ref = z["grad_" + name.replace(".", "_")]
        np.testing.assert_allclose(prm.grad.cpu().numpy(), ref, rtol=2e-3, atol=2e-5 * float(np.abs(ref).max()))
