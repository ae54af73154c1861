"""Parity at BASELINE.json's full sizes against the C restatement of the oracle (oracle/ballenv_oracle.c).

config 2: 4096 envs, WINDOW=5, 200 steps, one launch per step (closed loop), default obstacles.
config 3: 65536 envs, WINDOW=10, dense moving obstacles, through the rollout kernel (one launch per chunk).
Every observation, done and flag byte is compared bit for bit at every step; rewards to 1e-5 relative
(fp32 storage of the fp64 reward)."""
import numpy as np
import pytest
import torch

from helpers import oracle_config

pytestmark = pytest.mark.gpu


def _env_config(cfg):
    from gym_ballenv_b200 import EnvConfig
    return EnvConfig(static_obstacles=cfg["static_obstacles"], dynamic_obstacles=cfg["dynamic_obstacles"],
                     obstacle_speed=cfg["obstacle_speed"], obs_goal_position=cfg["obs_goal_position"],
                     time_step_for_change=cfg["time_step_for_change"], rd_th_obs=cfg["rd_th_obs"],
                     static_penalty=cfg["static_penalty"], dynamic_penalty=cfg["dynamic_penalty"])


def _check_state(env, c):
    st, ref = env.get_state(), c.state()
    ks = env.config.static_obstacles
    assert np.array_equal(st["agent_x"].cpu().numpy(), ref["agent"][:, 0].astype(np.float32))
    assert np.array_equal(st["agent_y"].cpu().numpy(), ref["agent"][:, 1].astype(np.float32))
    assert np.array_equal(st["goal_x"].cpu().numpy(), ref["goal"][:, 0].astype(np.float32))
    assert np.array_equal(st["dist"].cpu().numpy(), ref["dist"])                      # fp64, same operations
    assert np.array_equal(st["total_distance"].cpu().numpy(), ref["total"])
    np.testing.assert_allclose(st["acc_reward"].cpu().numpy(), ref["acc"], rtol=1e-12, atol=1e-12)
    assert np.array_equal(st["ep_len"].cpu().numpy(), ref["ep_len"])
    assert np.array_equal(st["episode"].cpu().numpy(), ref["episode"])
    assert np.array_equal(st["static_x"].cpu().numpy().T, ref["obstacles"][:, :ks, 0].astype(np.float32))
    assert np.array_equal(st["dynamic_x"].cpu().numpy().T, ref["obstacles"][:, ks:, 0].astype(np.float32))
    assert np.array_equal(st["dynamic_y"].cpu().numpy().T, ref["obstacles"][:, ks:, 1].astype(np.float32))
    assert np.array_equal(st["dynamic_goal"].cpu().numpy().T, ref["dyn_goal"])
    assert np.array_equal(st["dynamic_counter"].cpu().numpy().T, ref["dyn_counter"])


def test_config2_4096_envs_w5_200_steps_closed_loop():
    from gym_ballenv_b200 import BallVecEnv
    from oracle.c_oracle import COracleVec
    from oracle.gen_golden import CFG_DEFAULT
    n, T, seed = 4096, 200, 2024
    env = BallVecEnv(n, window=5, config=_env_config(CFG_DEFAULT), seed=seed, max_episode_steps=60)
    c = COracleVec(oracle_config(CFG_DEFAULT, 5, 60), seed, n)
    obs = env.reset()
    c.reset()
    assert np.array_equal(obs.cpu().numpy(), c.observe())
    g = torch.Generator().manual_seed(8)
    for t in range(T):
        a = torch.randint(0, 9, (n,), generator=g)
        obs, rew, done, info = env.step(a.cuda())
        r, d, f = c.step(a.numpy())
        assert np.array_equal(obs.cpu().numpy(), c.observe()), t
        assert np.array_equal(done.cpu().numpy(), d), t
        assert np.array_equal(info["flags"].cpu().numpy(), f), t
        np.testing.assert_allclose(rew.cpu().numpy(), r, rtol=1e-5, atol=0)
    _check_state(env, c)
    st = env.stats()
    for k, v in c.stats.items():
        assert st[k] == pytest.approx(v, rel=1e-9), k
    assert st["episodes"] > 4096 and env.error_flags() == 0
    env.close()


def test_config3_65536_envs_w10_dense_rollout_kernel():
    from gym_ballenv_b200 import BallVecEnv
    from oracle.c_oracle import COracleVec
    from oracle.gen_golden import CFG_DENSE
    n, T, seed = 65536, 24, 7
    env = BallVecEnv(n, window=10, config=_env_config(CFG_DENSE), seed=seed, max_episode_steps=10)
    c = COracleVec(oracle_config(CFG_DENSE, 10, 10), seed, n)
    obs = env.reset()
    c.reset()
    assert np.array_equal(obs.cpu().numpy(), c.observe())
    g = torch.Generator().manual_seed(9)
    a = torch.randint(0, 9, (T, n), generator=g)
    l0 = env.launch_count
    obs, rew, done = env.step_many(a.cuda(), keep_all_obs=True)
    assert env.launch_count - l0 == 1                       # the whole chunk is one launch of the rollout kernel
    obs, rew, done = obs.cpu().numpy(), rew.cpu().numpy(), done.cpu().numpy()
    for t in range(T):
        r, d, f = c.step(a[t].numpy())
        assert np.array_equal(obs[t], c.observe()), t
        assert np.array_equal(done[t], d), t
        np.testing.assert_allclose(rew[t], r, rtol=1e-5, atol=0)
    assert np.array_equal(env.state_views["flags"].cpu().numpy(), f)
    _check_state(env, c)
    st = env.stats()
    for k, v in c.stats.items():
        assert st[k] == pytest.approx(v, rel=1e-9), k
    assert st["episodes"] >= 2 * n and env.error_flags() == 0
    env.close()


def test_round_trip_properties_at_full_size():
    """Size-independent properties at 65536 envs x 200 steps: observations are 0/1 with exactly one goal-quadrant
    bit; window rows 0 and 1 are equal (the reference's row-offset quirk); done <=> a flag is set; the agent's own
    cell is occupied iff the hit flag is set (no auto-reset, so the observation is of the same state as the flags);
    episode bookkeeping adds up."""
    from gym_ballenv_b200 import BallVecEnv, EnvConfig
    n, T, W = 65536, 200, 10
    env = BallVecEnv(n, window=W, config=EnvConfig.dense_moving(), seed=3, auto_reset=False, max_episode_steps=0)
    env.reset()
    g = torch.Generator(device="cuda").manual_seed(1)
    a = torch.randint(0, 9, (T, n), generator=g, device="cuda")
    obs, rew, done = env.step_many(a, keep_all_obs=False)
    flags = env.state_views["flags"]
    assert torch.all((obs == 0) | (obs == 1))
    assert torch.all(obs[:, :4].sum(1) == 1)
    grid = obs[:, 4:].view(n, W, W)
    assert torch.equal(grid[:, 0], grid[:, 1])
    assert torch.equal(done[-1], flags != 0)
    own = grid[:, W // 2 + 1, W // 2] > 0          # row r samples y offset r - 1 - h: the agent's own cell
    assert torch.equal(own, (flags & 2) != 0)
    st = env.stats()
    assert st["steps"] == n * T
    assert env.error_flags() == 0
    env.close()


@pytest.mark.parametrize("ks,kd,w,n", [(40, 12, 5, 100), (0, 5, 5, 70), (3, 0, 7, 33), (70, 40, 10, 64)])
def test_unusual_obstacle_counts_against_the_c_oracle(ks, kd, w, n):
    """More than 8 quads per environment (the slots beyond the 256 register-held ones go through global memory
    every step, dynamic ones included), no static or no dynamic obstacles: per-step launches and the rollout
    kernel, with auto-resets, against the C oracle."""
    from gym_ballenv_b200 import BallVecEnv, EnvConfig
    from oracle.ballenv_oracle import OracleConfig
    from oracle.c_oracle import COracleVec
    goals = [(13 * i % 480 + 7, 29 * i % 440 + 30) for i in range(max(kd, 2))]
    speeds = [1 + (j % 3) for j in range(kd)]
    cfg = EnvConfig(static_obstacles=ks, dynamic_obstacles=kd, obstacle_speed=speeds,
                    obs_goal_position=["%d,%d" % g for g in goals] if kd else (), time_step_for_change=7)
    ocfg = OracleConfig(window=w, n_static=ks, n_dynamic=kd, speeds=speeds, goals=goals, change_step=7,
                        max_episode_steps=11)
    seed, T = 17, 30
    env = BallVecEnv(n, window=w, config=cfg, seed=seed, max_episode_steps=11)
    c = COracleVec(ocfg, seed, n)
    obs = env.reset()
    c.reset()
    assert np.array_equal(obs.cpu().numpy(), c.observe())
    g = torch.Generator().manual_seed(4)
    for t in range(10):                                   # one launch per step
        a = torch.randint(0, 9, (n,), generator=g)
        obs, rew, done, info = env.step(a.cuda())
        r, d, f = c.step(a.numpy())
        assert np.array_equal(obs.cpu().numpy(), c.observe()), t
        assert np.array_equal(done.cpu().numpy(), d) and np.array_equal(info["flags"].cpu().numpy(), f), t
        np.testing.assert_allclose(rew.cpu().numpy(), r, rtol=1e-5, atol=0)
    a = torch.randint(0, 9, (T, n), generator=g)          # one launch for T steps
    obs, rew, done = env.step_many(a.cuda(), keep_all_obs=True)
    obs, rew, done = obs.cpu().numpy(), rew.cpu().numpy(), done.cpu().numpy()
    for t in range(T):
        r, d, f = c.step(a[t].numpy())
        assert np.array_equal(obs[t], c.observe()), t
        assert np.array_equal(done[t], d), t
        np.testing.assert_allclose(rew[t], r, rtol=1e-5, atol=0)
    _check_state(env, c)
    assert env.stats()["episodes"] == c.stats["episodes"] > 0 and env.error_flags() == 0
    env.close()


@pytest.mark.parametrize("parity", [False, True])
def test_duplicate_obstacle_goals_against_the_c_oracle(parity):
    """Repeated obstacle goals (4 goals, two of them the same point; change step 5): the kernels' generic goal-change
    branch - candidates are the goals whose position differs from the current one (ballenv_env.py:351-352) - at 2048
    environments x 150 steps against oracle/ballenv_oracle.c, every observation / done / flag, then the state (goal
    indices through their canonical position: the reference only holds positions)."""
    from gym_ballenv_b200 import BallVecEnv
    from helpers import canonical_goal_index
    from oracle.c_oracle import COracleVec
    from oracle.gen_golden import CFG_DUPGOALS
    n, T, seed = 2048, 150, 77
    canon = canonical_goal_index(CFG_DUPGOALS)
    env = BallVecEnv(n, window=5, config=_env_config(CFG_DUPGOALS), seed=seed, max_episode_steps=37, parity=parity)
    assert env.kernel_variant(1) == "generic"        # repeated goals are outside the production specialisations
    c = COracleVec(oracle_config(CFG_DUPGOALS, 5, 37), seed, n)
    obs = env.reset()
    c.reset()
    assert np.array_equal(obs.cpu().numpy(), c.observe())
    g = torch.Generator().manual_seed(4)
    for t in range(T):
        a = torch.randint(0, 9, (n,), generator=g)
        obs, rew, done, info = env.step(a.cuda())
        r, d, f = c.step(a.numpy())
        assert np.array_equal(obs.cpu().numpy(), c.observe()), t
        assert np.array_equal(done.cpu().numpy(), d), t
        assert np.array_equal(info["flags"].cpu().numpy(), f), t
        np.testing.assert_allclose(rew.cpu().numpy(), r, rtol=1e-5, atol=0)
    st, ref = env.get_state(), c.state()
    assert np.array_equal(st["dynamic_x"].cpu().numpy().T.astype(np.float64), ref["obstacles"][:, 2:, 0])
    assert np.array_equal(st["dynamic_y"].cpu().numpy().T.astype(np.float64), ref["obstacles"][:, 2:, 1])
    assert np.array_equal(canon[st["dynamic_goal"].cpu().numpy().T], canon[ref["dyn_goal"]])
    assert np.array_equal(st["dynamic_counter"].cpu().numpy().T, ref["dyn_counter"])
    assert env.stats()["episodes"] > n and env.error_flags() == 0
    env.close()
