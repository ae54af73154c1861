"""The policy-in-the-loop rollout as one launch (ballenv_rollout_policy / FusedRollout; BASELINE config 5,
examples/ball_cnn_ac3.py:553-613).

What is pinned: (1) the environment side is the rollout kernel itself - replaying the actions the fused launch took
through ballenv_step_many on a twin environment gives bit-identical observations, rewards and dones; (2) the actions
are Categorical(policy(obs)) by inverse CDF of the documented draw: recomputed here from the torch policy's own
probabilities and the numpy Philox word (oracle/draws.py), the emitted action is the one u selects (float32 summation
order differs between the kernel and cuBLAS, so a draw within 1e-5 of a CDF step may go either way); (3) greedy mode is
the argmax of the torch policy; (4) the training loop on top of it runs."""
import os

import numpy as np
import pytest

from oracle import draws as D

SEED = 11


def _make(n, ks=13, kd=5, window=5, lanes=None, g0=0):
    import torch
    from gym_ballenv_b200 import BallVecEnv, EnvConfig
    if kd == 24:
        goals = ['%d,%d' % (x, y) for y in (100, 200, 300, 400) for x in (50, 130, 210, 290, 370, 450)]
        cfg = EnvConfig(static_obstacles=8, dynamic_obstacles=24, obstacle_speed=[1] * 24, obs_goal_position=goals)
    else:
        cfg = EnvConfig()
    old = os.environ.get("BALLENV_LEAN_G")
    if lanes is not None:
        os.environ["BALLENV_LEAN_G"] = str(lanes)      # read when the handle is created
    try:
        env = BallVecEnv(n, window=window, config=cfg, seed=SEED, device="cuda:0", max_episode_steps=25,
                         global_env_offset=g0)
    finally:
        if lanes is not None:
            if old is None:
                del os.environ["BALLENV_LEAN_G"]
            else:
                os.environ["BALLENV_LEAN_G"] = old
    return env, torch


def _policy(torch, window=5, scale=3.0):
    from gym_ballenv_b200.a2c import Policy
    torch.manual_seed(5)
    pol = Policy(window).to("cuda:0")
    with torch.no_grad():      # sharper than the default initialisation: the actions really depend on the observation
        pol.action_head.weight.mul_(scale)
        pol.fc1.weight.mul_(scale)
    return pol


def _fused(env, torch, pol, T, greedy=False):
    n, row = env.num_envs, env.obs_row
    first = env.observe().clone()
    obs = torch.zeros((T, n, row), dtype=torch.float32, device=env.device)
    act = torch.zeros((T, n), dtype=torch.int64, device=env.device)
    rew = torch.zeros((T, n), dtype=torch.float32, device=env.device)
    done = torch.zeros((T, n), dtype=torch.uint8, device=env.device)
    env.rollout_policy(pol, T, first, obs, act, rew, done, greedy=greedy)
    torch.cuda.synchronize()
    assert env.error_flags() == 0
    return first, obs, act, rew, done


@pytest.mark.gpu
@pytest.mark.parametrize("lanes", [1, 2])
@pytest.mark.parametrize("kd", [5, 24])
def test_fused_rollout_is_the_rollout_kernel_under_the_policys_actions(lanes, kd):
    n, T, g0 = 2048 + 17, 60, 1000
    env, torch = _make(n, kd=kd, lanes=lanes, g0=g0)
    twin, _ = _make(n, kd=kd, g0=g0)
    env.reset()
    twin.reset()
    pol = _policy(torch)
    tick0 = env.get_state()["tick"].cpu().numpy().astype(np.uint64)
    first, obs, act, rew, done = _fused(env, torch, pol, T)
    assert int(act.min()) >= 0 and int(act.max()) <= 8 and len(torch.unique(act)) == 9
    assert int(done.sum()) > n        # episodes end (time limit 25) and restart inside the launch
    # (1) the environment under these actions
    o2, r2, d2 = twin.step_many(act, keep_all_obs=True)
    assert torch.equal(o2, obs) and torch.equal(r2, rew) and torch.equal(d2.to(torch.uint8), done)
    sa, sb = env.get_state(), twin.get_state()
    for k in sa:
        assert torch.equal(sa[k], sb[k]), k
    # (2) the actions under the policy: u of the documented draw against the torch policy's own CDF
    seen = torch.cat([first.unsqueeze(0), obs[:-1]], 0)
    with torch.no_grad():
        probs, _ = pol(seen.reshape(T * n, -1))
    cdf = torch.cumsum(probs.double(), -1).cpu().numpy().reshape(T, n, 9)
    genv = (np.arange(n, dtype=np.uint64) + np.uint64(g0))[None, :]
    tick = tick0[None, :] + np.arange(T, dtype=np.uint64)[:, None]
    w = D.philox4x32_10_np(np.broadcast_to(genv, (T, n)), tick, 0, D.STREAM_ACTION, SEED & 0xffffffff, SEED >> 32)[0]
    u = (w >> np.uint64(8)).astype(np.float64) / 16777216.0
    a = act.cpu().numpy()
    want = np.minimum((u[..., None] >= cdf[..., :8]).sum(-1), 8)
    edge = np.abs(u[..., None] - cdf[..., :8]).min(-1) < 1e-5
    assert ((want == a) | edge).all(), int(((want != a) & ~edge).sum())
    assert (want == a).mean() > 0.9999
    env.close()
    twin.close()


@pytest.mark.gpu
def test_fused_greedy_is_the_argmax_of_the_torch_policy():
    n, T = 1024, 40
    env, torch = _make(n)
    env.reset()
    pol = _policy(torch)
    first, obs, act, _, _ = _fused(env, torch, pol, T, greedy=True)
    seen = torch.cat([first.unsqueeze(0), obs[:-1]], 0)
    with torch.no_grad():
        probs, _ = pol(seen.reshape(T * n, -1))
    top2 = probs.topk(2, -1).values
    tie = (top2[:, 0] - top2[:, 1]) < 1e-5
    ok = (probs.argmax(-1) == act.reshape(-1)) | tie
    assert bool(ok.all())
    env.close()


@pytest.mark.gpu
def test_fused_rollout_does_not_depend_on_the_sharding():
    """Environments [256, 512) of a 1024-environment job stepped alone take the same actions (draws are keyed by the
    global environment id)."""
    T = 30
    env, torch = _make(1024)
    part, _ = _make(256, g0=256)
    env.reset()
    part.reset()
    pol = _policy(torch)
    _, obs, act, rew, _ = _fused(env, torch, pol, T)
    _, obs_p, act_p, rew_p, _ = _fused(part, torch, pol, T)
    assert torch.equal(act[:, 256:512], act_p) and torch.equal(obs[:, 256:512], obs_p) and torch.equal(rew[:, 256:512], rew_p)
    env.close()
    part.close()


@pytest.mark.gpu
def test_fused_rollout_needs_a_configuration_it_is_built_for():
    from gym_ballenv_b200 import BallVecEnv
    from gym_ballenv_b200._lib import BallenvError
    import torch
    env = BallVecEnv(64, window=5, device="cuda:0", parity=True)       # fp64 parity mode: no lean kernel
    env.reset()
    pol = _policy(torch)
    with pytest.raises((BallenvError, ValueError)):
        _fused(env, torch, pol, 4)
    env.close()


@pytest.mark.gpu
@pytest.mark.parametrize("kd", [5, 24])
def test_fused_rollout_window_10(kd):
    """WINDOW = 10 (104 inputs, 208 hidden units - the reference's other shipped shape): same two checks."""
    n, T = 1500, 30
    env, torch = _make(n, kd=kd, window=10)
    twin, _ = _make(n, kd=kd, window=10)
    env.reset()
    twin.reset()
    pol = _policy(torch, window=10, scale=2.0)
    first, obs, act, rew, done = _fused(env, torch, pol, T, greedy=True)
    o2, r2, d2 = twin.step_many(act, keep_all_obs=True)
    assert torch.equal(o2, obs) and torch.equal(r2, rew) and torch.equal(d2.to(torch.uint8), done)
    seen = torch.cat([first.unsqueeze(0), obs[:-1]], 0)
    with torch.no_grad():
        probs, _ = pol(seen.reshape(T * n, -1))
    top2 = probs.topk(2, -1).values
    assert bool(((probs.argmax(-1) == act.reshape(-1)) | ((top2[:, 0] - top2[:, 1]) < 1e-5)).all())
    assert len(torch.unique(act)) >= 5
    env.close()
    twin.close()


@pytest.mark.gpu
def test_train_fused_updates_the_policy():
    from gym_ballenv_b200.a2c import FusedRollout, Policy, train_fused
    env, torch = _make(4096)
    torch.manual_seed(0)
    pol = Policy(5).to("cuda:0")
    before = [p.detach().clone() for p in pol.parameters()]
    losses = []
    train_fused(env, pol, iterations=4, n_steps=16, log=lambda it, loss, batch: losses.append(float(loss.detach())))
    assert len(losses) == 4 and all(np.isfinite(losses))
    assert any(not torch.equal(a, b) for a, b in zip(before, pol.parameters()))
    # the evaluate() of the stored pairs is the policy's own forward pass
    roll = FusedRollout(env, pol, 8)
    raw = roll.run()
    ev = roll.evaluate(raw)
    with torch.no_grad():
        probs, value = pol(raw["obs"][:8].reshape(8 * 4096, -1))
    lp = torch.log(probs.gather(-1, raw["action"].reshape(-1, 1)).squeeze(-1)).view(8, 4096)
    assert torch.allclose(ev["log_prob"], lp) and torch.allclose(ev["value"], value.view(8, 4096))
    assert env.error_flags() == 0
    env.close()


@pytest.mark.gpu
def test_discounted_returns_kernel_is_the_loop():
    import torch
    from gym_ballenv_b200.a2c import discounted_returns
    g = torch.Generator(device="cuda:0").manual_seed(3)
    T, n = 37, 5000
    reward = torch.randn((T, n), device="cuda:0", generator=g)
    done = torch.rand((T, n), device="cuda:0", generator=g) < 0.1
    boot = torch.randn((n,), device="cuda:0", generator=g)
    for b in (None, boot):
        for d in (done, done.to(torch.uint8)):
            got = discounted_returns(reward, d, 0.99, b)
            R = torch.zeros_like(reward[0]) if b is None else b          # the tensor-expression form (a2c.py)
            want = torch.empty_like(reward)
            for t in range(T - 1, -1, -1):
                R = reward[t] + 0.99 * R * (~done[t]).to(reward.dtype)
                want[t] = R
            assert torch.equal(got, want)


@pytest.mark.gpu
def test_graphed_trainer_is_the_eager_iteration():
    """The whole iteration replayed as one CUDA graph gives what the same steps give issued one by one: same rollouts
    (deterministic draws), same losses, same weights after the same number of iterations."""
    from gym_ballenv_b200.a2c import FusedRollout, GraphedTrainer, Policy, a2c_loss
    results = []
    for graphed in (True, False):
        env, torch = _make(2048)
        env.reset()
        torch.manual_seed(0)
        pol = Policy(5).to("cuda:0")
        losses = []
        if graphed:
            tr = GraphedTrainer(env, pol, n_steps=8, fused_update=False)
            tr.step()                       # three eager iterations + the capture
            for _ in range(4):
                losses.append(float(tr.step()))
        else:
            opt = torch.optim.Adam(pol.parameters(), lr=1e-3, capturable=True)
            roll = FusedRollout(env, pol, 8)
            for it in range(7):
                raw = roll.run()
                batch = roll.evaluate(raw)
                with torch.no_grad():
                    _, v_last = pol(raw["obs"][8])
                loss = a2c_loss(batch, 0.99, bootstrap=v_last.squeeze(-1))
                opt.zero_grad(set_to_none=True)
                loss.backward()
                opt.step()
                if it >= 3:
                    losses.append(float(loss))
        results.append((losses, [p.detach().clone() for p in pol.parameters()], env.get_state()))
        assert env.error_flags() == 0
        env.close()
    (la, pa, sa), (lb, pb, sb) = results
    assert np.allclose(la, lb, rtol=1e-4), (la, lb)
    for a, b in zip(pa, pb):
        assert torch.allclose(a, b, rtol=1e-3, atol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("window,n,T", [(5, 3000, 13), (10, 700, 9), (5, 100, 1)])
def test_fused_update_is_autograd(window, n, T):
    """ballenv_a2c_grads against torch.autograd on the same batch: the loss of a2c_loss and the gradient of every parameter
    (float32 sums of up to 39 000 terms in another order: 2e-4 relative to the gradient's scale)."""
    import torch
    from gym_ballenv_b200.a2c import FusedRollout, FusedUpdate, a2c_loss, normalised_returns
    env, torch = _make(n, window=window)
    env.reset()
    pol = _policy(torch, window=window, scale=1.5)
    roll = FusedRollout(env, pol, T)
    raw = roll.run()
    with torch.no_grad():
        _, v_last = pol(raw["obs"][T])
    batch = roll.evaluate(raw)
    loss = a2c_loss(batch, 0.99, bootstrap=v_last.squeeze(-1))
    pol.zero_grad(set_to_none=True)
    loss.backward()
    want = [p.grad.detach().clone() for p in pol.parameters()]
    pol.zero_grad(set_to_none=True)
    upd = FusedUpdate(pol, T * n)
    returns = normalised_returns(raw["reward"], roll.done, 0.99, bootstrap=v_last.squeeze(-1))
    got_loss = upd.grads(raw["obs"][:T], raw["action"], returns)
    assert torch.allclose(got_loss, loss.detach(), rtol=2e-4), (float(got_loss), float(loss))
    for (name, p), w in zip(pol.named_parameters(), want):
        scale = float(w.abs().max())
        assert scale > 0, name
        assert float((p.grad - w).abs().max()) <= 2e-4 * scale, (name, float((p.grad - w).abs().max()), scale)
    # deterministic: a second call gives the same bits
    g1 = [p.grad.clone() for p in pol.parameters()]
    upd.grads(raw["obs"][:T], raw["action"], returns)
    assert all(torch.equal(a, p.grad) for a, p in zip(g1, pol.parameters()))
    env.close()


@pytest.mark.gpu
def test_graphed_trainer_with_the_fused_update_learns_like_autograd():
    """Whole iterations: rollout launch + FusedUpdate + Adam replayed as one graph against the autograd update - same
    rollouts (deterministic draws), losses and weights equal to float32 rounding after the same number of iterations."""
    from gym_ballenv_b200.a2c import GraphedTrainer, Policy
    out = []
    for fused in (True, False):
        env, torch = _make(2048)
        env.reset()
        torch.manual_seed(0)
        pol = Policy(5).to("cuda:0")
        tr = GraphedTrainer(env, pol, n_steps=8, fused_update=fused)
        losses = [float(tr.step()) for _ in range(5)]
        out.append((losses, [p.detach().clone() for p in pol.parameters()]))
        assert env.error_flags() == 0
        env.close()
    (la, pa), (lb, pb) = out
    assert np.allclose(la, lb, rtol=1e-3), (la, lb)
    for a, b in zip(pa, pb):
        assert torch.allclose(a, b, rtol=1e-2, atol=1e-4)


@pytest.mark.gpu
def test_a2c_entry_points_reject_bad_arguments():
    import ctypes as C
    import torch
    from gym_ballenv_b200 import _lib as L
    assert L.LIB.ballenv_a2c_workspace_bytes(29, 130, 100) < 0          # hidden not a multiple of 4
    assert L.LIB.ballenv_a2c_workspace_bytes(29, 512, 100) < 0          # too wide
    assert L.LIB.ballenv_a2c_workspace_bytes(29, 128, 100) > 0
    u = L.BallenvA2CUpdate()
    u.n_inputs, u.hidden = 29, 128
    x = torch.zeros(64, device="cuda:0")
    assert L.LIB.ballenv_a2c_grads(C.byref(u), C.c_void_p(x.data_ptr()), C.c_void_p(x.data_ptr()), C.c_void_p(x.data_ptr()),
                                   4, C.c_void_p(x.data_ptr()), 256, None) == -1    # NULL parameter pointers
    assert b"NULL" in L.LIB.ballenv_last_error()
    assert L.LIB.ballenv_discounted_returns(None, None, None, C.c_float(0.9), 1, 1, None, None) == -1


@pytest.mark.gpu
@pytest.mark.parametrize("n,T", [(1, 1), (1, 7), (63, 2), (65, 1)])
def test_fused_rollout_and_update_at_tiny_sizes(n, T):
    """One environment, one step, ragged warps: the fused rollout against the rollout kernel, the fused update against
    autograd."""
    from gym_ballenv_b200.a2c import FusedRollout, FusedUpdate, a2c_loss, normalised_returns
    env, torch = _make(n)
    twin, _ = _make(n)
    env.reset()
    twin.reset()
    pol = _policy(torch)
    roll = FusedRollout(env, pol, T)
    raw = roll.run()
    o2, r2, d2 = twin.step_many(raw["action"], keep_all_obs=True)
    assert torch.equal(o2, raw["obs"][1:]) and torch.equal(r2, raw["reward"]) and torch.equal(d2, raw["done"])
    if n * T > 1:      # (the normalisation needs two samples)
        with torch.no_grad():
            _, v_last = pol(raw["obs"][T])
        loss = a2c_loss(roll.evaluate(raw), 0.99, bootstrap=v_last.squeeze(-1))
        pol.zero_grad(set_to_none=True)
        loss.backward()
        want = [p.grad.detach().clone() for p in pol.parameters()]
        pol.zero_grad(set_to_none=True)
        upd = FusedUpdate(pol, T * n)
        got = upd.grads(raw["obs"][:T], raw["action"],
                        normalised_returns(raw["reward"], roll.done, 0.99, bootstrap=v_last.squeeze(-1)))
        assert torch.allclose(got, loss.detach(), rtol=1e-4, atol=1e-5)
        for p, w in zip(pol.parameters(), want):
            assert float((p.grad - w).abs().max()) <= 2e-4 * max(float(w.abs().max()), 1e-3)
    env.close()
    twin.close()


@pytest.mark.gpu
@pytest.mark.parametrize("window,ks,kd,lanes", [(5, 10, 7, 2), (10, 3, 2, 1), (5, 20, 30, 2), (5, 0, 1, 2)])
def test_fused_rollout_with_run_time_obstacle_counts(window, ks, kd, lanes, monkeypatch):
    """Configurations without a tuned instance: the policy in the loop of the run-time-count kernels - the environment
    side against ballenv_step_many on a twin, greedy actions against torch's argmax."""
    import torch
    from gym_ballenv_b200 import BallVecEnv, EnvConfig
    goals = ["%d,%d" % (30 + (53 * i) % 440, 25 + (97 * i) % 450) for i in range(max(kd, 2) + 1)]
    cfg = EnvConfig(static_obstacles=ks, dynamic_obstacles=kd, obstacle_speed=[1 + (j % 2) for j in range(kd)],
                    obs_goal_position=goals, time_step_for_change=9)
    n, T = 700, 40
    monkeypatch.setenv("BALLENV_LEAN_G", str(lanes))
    env = BallVecEnv(n, window=window, config=cfg, seed=SEED, device="cuda:0", max_episode_steps=21)
    monkeypatch.delenv("BALLENV_LEAN_G")
    twin = BallVecEnv(n, window=window, config=cfg, seed=SEED, device="cuda:0", max_episode_steps=21)
    env.reset()
    twin.reset()
    pol = _policy(torch, window=window, scale=2.0)
    for greedy in (False, True):
        first, obs, act, rew, done = _fused(env, torch, pol, T, greedy=greedy)
        o2, r2, d2 = twin.step_many(act, keep_all_obs=True)
        assert torch.equal(o2, obs) and torch.equal(r2, rew) and torch.equal(d2.to(torch.uint8), done)
        if greedy:
            seen = torch.cat([first.unsqueeze(0), obs[:-1]], 0)
            with torch.no_grad():
                probs, _ = pol(seen.reshape(T * n, -1))
            top2 = probs.topk(2, -1).values
            assert bool(((probs.argmax(-1) == act.reshape(-1)) | ((top2[:, 0] - top2[:, 1]) < 1e-5)).all())
    sa, sb = env.get_state(), twin.get_state()
    for k in sa:
        assert torch.equal(sa[k], sb[k]), k
    assert int(done.sum()) > 0
    env.close()
    twin.close()


@pytest.mark.gpu
@pytest.mark.parametrize("window,lanes", [(5, 2), (5, 1), (10, 2)])
def test_rollout_policy_out_is_the_policys_forward_pass(window, lanes):
    """policy_out of the fused rollout = the torch policy's probabilities and values on the observations it acted on, and
    the update fed with it gives the gradients of the update that recomputes the forward pass."""
    from gym_ballenv_b200.a2c import FusedRollout, FusedUpdate, normalised_returns
    n, T = 1200, 12
    env, torch = _make(n, window=window, lanes=lanes)
    env.reset()
    pol = _policy(torch, window=window, scale=1.5)
    roll = FusedRollout(env, pol, T, keep_policy_out=True)
    raw = roll.run()
    with torch.no_grad():
        probs, value = pol(raw["obs"][:T].reshape(T * n, -1))
        _, v_last = pol(raw["obs"][T])
    po = raw["policy_out"].reshape(T * n, 10)
    assert torch.allclose(po[:, :9], probs, rtol=1e-4, atol=1e-6)
    assert torch.allclose(po[:, 9], value.squeeze(-1), rtol=1e-4, atol=1e-5)
    returns = normalised_returns(raw["reward"], roll.done, 0.99, bootstrap=v_last.squeeze(-1))
    upd = FusedUpdate(pol, T * n)
    l1 = upd.grads(raw["obs"][:T], raw["action"], returns).clone()
    g1 = [p.grad.clone() for p in pol.parameters()]
    l2 = upd.grads(raw["obs"][:T], raw["action"], returns, policy_out=raw["policy_out"]).clone()
    assert torch.allclose(l1, l2, rtol=1e-5)
    for a, p in zip(g1, pol.parameters()):
        assert float((a - p.grad).abs().max()) <= 1e-4 * max(float(a.abs().max()), 1e-6)
    env.close()
