"""GPU parity: the CUDA path (through the C ABI, via gym_ballenv_b200) against
 (a) the golden fixtures recorded from the reference's own code, and
 (b) the CPU oracle on the same seeded inputs.
Integer/flag/grid results must be bit-exact; rewards agree to 1e-5 relative in fp32 mode (they are computed
in fp64 and rounded once, so the observed error is ~6e-8) and exactly in fp64 parity mode."""
import numpy as np
import pytest

from helpers import load_golden, oracle_config, parse_goals, tapes_from_golden

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def _env_config(cfg):
    from gym_ballenv_b200 import EnvConfig
    return EnvConfig(static_obstacles=cfg["static_obstacles"], dynamic_obstacles=cfg["dynamic_obstacles"],
                     obstacle_speed=cfg["obstacle_speed"], obs_goal_position=cfg["obs_goal_position"],
                     time_step_for_change=cfg["time_step_for_change"], rd_th_obs=cfg["rd_th_obs"],
                     static_penalty=cfg["static_penalty"], dynamic_penalty=cfg["dynamic_penalty"])


def _rows_from_obs(obs, w):
    """obs [N, 4 + w*w] (any numeric dtype) -> (quadrant [N], rows [N, w] bitmasks)."""
    o = obs.detach().cpu().numpy().astype(np.int64)
    assert set(np.unique(o)) <= {0, 1}
    assert np.all(o[:, :4].sum(1) == 1)
    q = o[:, :4].argmax(1)
    grid = o[:, 4:].reshape(-1, w, w)
    rows = (grid << np.arange(w)[None, None, :]).sum(2)
    return q, rows


def _check_state(env, z, prefix, idx, ks, kd, exact_dist=True):
    s = env.get_state()
    n = env.num_envs
    agent = np.stack([s["agent_x"].cpu().numpy(), s["agent_y"].cpu().numpy()], 1)
    goal = np.stack([s["goal_x"].cpu().numpy(), s["goal_y"].cpu().numpy()], 1)
    exp_agent = z[prefix + "agent"][idx] if idx is not None else z[prefix + "agent"]
    exp_goal = z[prefix + "goal"][idx] if idx is not None else z[prefix + "goal"]
    assert np.array_equal(agent, exp_agent.astype(np.float64))
    assert np.array_equal(goal, exp_goal.astype(np.float64))
    obst = np.concatenate([np.stack([s["static_x"].cpu().numpy(), s["static_y"].cpu().numpy()], 2),
                           np.stack([s["dynamic_x"].cpu().numpy(), s["dynamic_y"].cpu().numpy()], 2)], 0)
    exp_obst = z[prefix + "obst"][idx] if idx is not None else z[prefix + "obst"]
    assert np.array_equal(obst.transpose(1, 0, 2), exp_obst.astype(np.float64))
    exp_dist = z[prefix + "dist"][idx] if idx is not None else z[prefix + "dist"]
    exp_total = z[prefix + "total_distance"][idx] if idx is not None else z[prefix + "total_distance"]
    assert np.array_equal(s["dist"].cpu().numpy(), exp_dist)
    assert np.array_equal(s["total_distance"].cpu().numpy(), exp_total)
    return s


@pytest.mark.parametrize("parity", [False, True])
@pytest.mark.parametrize("name", ["rollout_philox_default", "rollout_philox_busy", "rollout_philox_dense",
                                  "rollout_mt_default", "rollout_philox_dupgoals", "rollout_checkpoint"])
def test_golden_rollout(name, parity):
    """Rollouts recorded by running the reference itself (oracle/gen_golden.py through oracle/ref_shim.py).
    rollout_philox_dupgoals has two equal obstacle goals: the goal-change step then takes the kernels' generic
    branch (move_obstacle with goals_distinct == 0: candidates are the goals whose POSITION differs,
    ballenv_env.py:351-352); a goal is compared by its canonical (first) index, as the reference only holds positions."""
    from gym_ballenv_b200 import BallVecEnv
    from helpers import canonical_goal_index
    z, meta = load_golden(name)
    cfg = meta["cfg"]
    canon = canonical_goal_index(cfg)
    ks, kd = cfg["static_obstacles"], cfg["dynamic_obstacles"]
    n, T, g0 = meta["n_envs"], meta["T"], meta["g0"]
    for w in meta["windows"]:
        env = BallVecEnv(n, window=w, config=_env_config(cfg), seed=meta["seed"], parity=parity,
                         max_episode_steps=meta["max_episode_steps"], global_env_offset=g0)
        if meta["mode"] == "mt":
            step, reset = tapes_from_golden(z, meta)
            env.set_draw_tape(step, reset, meta["tape_attempts"])
        obs = env.reset()
        _check_state(env, z, "init_", None, ks, kd)
        q, rows = _rows_from_obs(obs, w)
        assert np.array_equal(q, z["init_quadrant"]) and np.array_equal(rows, z["init_rows%d" % w])
        for t in range(T):
            a = torch.from_numpy(z["rec_actions"][t].astype(np.int64)).cuda()
            obs, rew, done, info = env.step(a)
            assert np.array_equal(done.cpu().numpy().astype(np.uint8), z["rec_done"][t]), (t, w)
            assert np.array_equal(info["flags"].cpu().numpy(), z["rec_flags"][t]), (t, w)
            if parity:
                assert np.array_equal(rew.cpu().numpy(), z["rec_reward"][t]), (t, w)
            else:
                np.testing.assert_allclose(rew.cpu().numpy(), z["rec_reward"][t], rtol=1e-5, atol=0)
            s = _check_state(env, z, "rec_", t, ks, kd)
            assert np.array_equal(s["acc_reward"].cpu().numpy(), z["rec_acc"][t]), (t, w)
            assert np.array_equal(s["ep_len"].cpu().numpy(), z["rec_ep_len"][t])
            assert np.array_equal(canon[s["dynamic_goal"].cpu().numpy().T], z["rec_dyn_goal"][t])
            assert np.array_equal(s["dynamic_counter"].cpu().numpy().T, z["rec_dyn_counter"][t])
            q, rows = _rows_from_obs(obs, w)
            assert np.array_equal(q, z["rec_quadrant"][t]), (t, w)
            assert np.array_equal(rows, z["rec_rows%d" % w][t]), (t, w)
        st = env.stats()
        for k, v in meta["stats"].items():
            assert st[k] == pytest.approx(v, rel=1e-9), k
        assert env.error_flags() == 0
        env.close()


@pytest.mark.parametrize("parity", [False, True])
def test_edge_cases(parity):
    from gym_ballenv_b200 import BallVecEnv
    z, meta = load_golden("edge_gym")
    cfg = meta["cfg"]
    n = len(meta["names"])
    for w in (5, 10):
        env = BallVecEnv(n, window=w, config=_env_config(cfg), parity=parity, auto_reset=False, max_episode_steps=0)
        env.reset()
        env.set_state(agent_x=z["in_agent"][:, 0], agent_y=z["in_agent"][:, 1], goal_x=z["in_goal"][:, 0],
                      goal_y=z["in_goal"][:, 1], dist=z["in_dist"], total_distance=z["in_total"],
                      acc_reward=np.full(n, meta["in_acc"]),
                      static_x=z["in_obst"][:, :2, 0].T, static_y=z["in_obst"][:, :2, 1].T,
                      dynamic_x=z["in_obst"][:, 2:, 0].T, dynamic_y=z["in_obst"][:, 2:, 1].T,
                      dynamic_goal=z["in_goal_idx"].T.astype(np.int32), dynamic_counter=z["in_counter"].T)
        env.set_draw_tape(z["words"][None])
        a = torch.from_numpy(z["action"]).cuda()
        obs, rew, done, info = env.step(a if parity else a.float())
        s = env.get_state()
        names = np.array(meta["names"])

        def same(got, exp, what):
            bad = np.nonzero(~np.all(np.asarray(got).reshape(n, -1) == np.asarray(exp).reshape(n, -1), axis=1))[0]
            assert len(bad) == 0, (what, list(names[bad]))

        same(np.stack([s["agent_x"].cpu().numpy(), s["agent_y"].cpu().numpy()], 1), z["out_agent"], "agent")
        same(s["dist"].cpu().numpy(), z["out_dist"], "dist")
        obst = np.concatenate([np.stack([s["static_x"].cpu().numpy(), s["static_y"].cpu().numpy()], 2),
                               np.stack([s["dynamic_x"].cpu().numpy(), s["dynamic_y"].cpu().numpy()], 2)], 0)
        same(obst.transpose(1, 0, 2), z["out_obst"], "obstacles")
        same(s["dynamic_goal"].cpu().numpy().T, z["out_goal_idx"], "goal idx")
        same(s["dynamic_counter"].cpu().numpy().T, z["out_counter"], "counter")
        same(done.cpu().numpy().astype(np.uint8), z["out_done"], "done")
        same(info["flags"].cpu().numpy(), z["out_flags"], "flags")
        if parity:
            same(rew.cpu().numpy(), z["out_reward"], "reward")
            same(s["acc_reward"].cpu().numpy(), z["out_acc"], "acc")
        else:
            np.testing.assert_allclose(rew.cpu().numpy(), z["out_reward"], rtol=1e-5, atol=0)
        q, rows = _rows_from_obs(obs, w)
        same(q, z["out_quadrant"], "quadrant")
        same(rows, z["out_rows%d" % w], "rows")
        env.close()


@pytest.mark.parametrize("parity", [False, True])
@pytest.mark.parametrize("w", [5, 10, 21])
def test_window_kat(w, parity):
    from gym_ballenv_b200 import BallVecEnv, EnvConfig
    z, meta = load_golden("window_kat")
    n, kmax = z["obst"].shape[0], z["obst"].shape[1]
    cfg = EnvConfig(static_obstacles=kmax, dynamic_obstacles=0, obstacle_speed=(), obs_goal_position=())
    env = BallVecEnv(n, window=w, config=cfg, parity=parity, auto_reset=False, max_episode_steps=0)
    env.set_state(agent_x=z["agent"][:, 0], agent_y=z["agent"][:, 1], goal_x=z["goal"][:, 0], goal_y=z["goal"][:, 1],
                  static_x=z["obst"][:, :, 0].T, static_y=z["obst"][:, :, 1].T)   # unused slots sit at (10000, 10000)
    q, rows = _rows_from_obs(env.observe(), w)
    assert np.array_equal(q, z["quadrant"])
    assert np.array_equal(rows, z["rows%d" % w])
    env.close()


@pytest.mark.parametrize("obs_dtype", ["f32", "u8", "bits"])
def test_obs_formats_agree(obs_dtype):
    from gym_ballenv_b200 import BallVecEnv
    dt = {"f32": torch.float32, "u8": torch.uint8, "bits": "bits"}[obs_dtype]
    n = 1000   # not a multiple of the block size: exercises the ragged tail of the coalesced store
    for w in (5, 10, 7):
        ref = BallVecEnv(n, window=w, seed=5, obs_dtype=torch.float32)
        env = BallVecEnv(n, window=w, seed=5, obs_dtype=dt)
        o_ref, o = ref.reset(), env.reset()
        g = torch.Generator().manual_seed(0)
        for t in range(30):
            a = torch.randint(0, 9, (n,), generator=g).cuda()
            o_ref = ref.step(a)[0]
            o = env.step(a)[0]
        if obs_dtype == "bits":
            nb = 4 + w * w
            words = o.cpu().numpy().astype(np.int64) & 0xffffffff
            bits = ((words[:, :, None] >> np.arange(32)[None, None, :]) & 1).reshape(n, -1)[:, :nb]
            assert np.array_equal(bits, o_ref.cpu().numpy().astype(np.int64))
        else:
            assert np.array_equal(o.cpu().numpy().astype(np.float32), o_ref.cpu().numpy())
        ref.close()
        env.close()


@pytest.mark.parametrize("w,cfgname", [(5, "default"), (10, "dense")])
def test_oracle_rollout(w, cfgname):
    """Config 2 shape at a size the Python oracle finishes in seconds; the 4096-env x 200-step case runs against
    the C oracle in test_gpu_parity_full.py."""
    from gym_ballenv_b200 import BallVecEnv, EnvConfig
    from oracle import draws as D
    from oracle.ballenv_oracle import OracleVec
    from oracle.gen_golden import CFG_DEFAULT, CFG_DENSE
    cfg = CFG_DEFAULT if cfgname == "default" else CFG_DENSE
    n, T, seed, g0 = 192, 40, 77, 123456
    env = BallVecEnv(n, window=w, config=_env_config(cfg), seed=seed, max_episode_steps=25, global_env_offset=g0)
    vec = OracleVec(oracle_config(cfg, w, 25), D.PhiloxDraws(seed), n, g0)
    obs = env.reset()
    vec.reset()
    assert np.array_equal(obs.cpu().numpy(), np.array(vec.observe(), dtype=np.float32))
    g = torch.Generator().manual_seed(3)
    for t in range(T):
        a = torch.randint(0, 9, (n,), generator=g)
        obs, rew, done, info = env.step(a.cuda())
        r, d, f = vec.step(a.tolist())
        assert np.array_equal(obs.cpu().numpy(), np.array(vec.observe(), dtype=np.float32)), t
        assert np.array_equal(done.cpu().numpy(), np.array(d)), t
        assert np.array_equal(info["flags"].cpu().numpy(), np.array(f, dtype=np.uint8)), t
        np.testing.assert_allclose(rew.cpu().numpy(), np.array(r), rtol=1e-5, atol=0)
    st = env.stats()
    for k, v in vec.stats.items():
        assert st[k] == pytest.approx(v, rel=1e-9), k
    env.close()


def test_shard_invariance():
    """Same global env id -> same trajectory regardless of how the envs are split into handles."""
    from gym_ballenv_b200 import BallVecEnv
    n, T = 512, 60
    whole = BallVecEnv(n, window=5, seed=9, max_episode_steps=20)
    parts = [BallVecEnv(n // 4, window=5, seed=9, max_episode_steps=20, global_env_offset=i * (n // 4)) for i in range(4)]
    o = whole.reset()
    op = torch.cat([p.reset() for p in parts])
    assert torch.equal(o, op)
    g = torch.Generator().manual_seed(2)
    for t in range(T):
        a = torch.randint(0, 9, (n,), generator=g).cuda()
        o, r, d, _ = whole.step(a)
        outs = [p.step(a[i * (n // 4):(i + 1) * (n // 4)]) for i, p in enumerate(parts)]
        assert torch.equal(o, torch.cat([x[0] for x in outs]))
        assert torch.equal(r, torch.cat([x[1] for x in outs]))
        assert torch.equal(d, torch.cat([x[2] for x in outs]))
    tot = whole.stats()
    summed = {k: sum(p.stats()[k] for p in parts) for k in tot}
    for k in tot:
        assert summed[k] == pytest.approx(tot[k], rel=1e-9)


def test_single_env_facade():
    """gym-style API (reset/step/state list) against the oracle's state list."""
    import gym_ballenv_b200 as gb
    from argparse import Namespace
    from oracle import draws as D
    from oracle.ballenv_oracle import OracleConfig, OracleEnv
    from oracle.gen_golden import CFG_DEFAULT
    args = Namespace(rd_th_agent=80, static_thresholds=[0, 0], dynamic_thresholds=[10, 10], **CFG_DEFAULT)
    env = gb.make('gymball-v0', seed=21)
    env.unwrapped.customize_environment(args)
    prep_state2, prep_state4 = gb.make_prep_state(env)
    ocfg = oracle_config(CFG_DEFAULT, 5, 0, False)
    ref = OracleEnv(ocfg, D.PhiloxDraws(21), 0)
    state = env.reset()
    assert list(state) == ref.reset()
    rng = np.random.RandomState(0)
    for t in range(50):
        obs = prep_state4(state, 5)
        assert obs.shape == (1, 29) and obs.dtype == torch.float32 and obs.is_cuda
        assert obs.cpu().numpy().reshape(-1).tolist() == ref.observe(5)
        a = gb.MOVE_LIST[rng.randint(9)]
        state, reward, done, info = env.step(a)
        r, d = ref.step(a)
        assert list(state) == ref.state() and reward == r and done == d
        assert env.unwrapped.total_reward_accumulated == ref.acc
        if done:
            break
    # arbitrary-state observation at another window size
    st = [(82, 82), (400, 490), 0.0, (100, 100)]
    o10 = prep_state4(st, 10).cpu().numpy().reshape(-1)
    from oracle.ballenv_oracle import window_obs
    assert o10.tolist() == window_obs(st[0], st[1], st[3:], 10)
    env.close()


def test_bad_action_flag_and_errors():
    from gym_ballenv_b200 import BallVecEnv, BallenvError, EnvConfig
    env = BallVecEnv(64, window=5)
    env.reset()
    a = torch.full((64,), 11, dtype=torch.int64, device="cuda")
    env.step(a)
    assert env.error_flags() & 1
    assert env.error_flags() == 0     # read-and-clear
    with pytest.raises(ValueError):
        env.step(torch.zeros(63, dtype=torch.int64, device="cuda"))
    with pytest.raises(BallenvError):
        BallVecEnv(8, window=40)
    with pytest.raises(BallenvError):   # a single distinct obstacle goal: the reference raises at the first change step
        BallVecEnv(8, config=EnvConfig(dynamic_obstacles=2, obstacle_speed=[1, 1], obs_goal_position=['5,5', '5,5']))
    env.close()


@pytest.mark.gpu
@pytest.mark.parametrize("no_lean", [False, True])
@pytest.mark.parametrize("w,cfgname,n", [(5, "default", 1000), (10, "dense", 777), (7, "default", 100)])
def test_fast_and_generic_kernels_agree(w, cfgname, n, no_lean, monkeypatch):
    """The production specialisation (ballenv_kernel<..., kFast=true>) and the generic kernel are the same
    function: bit-identical observations, rewards, flags and state over a rollout with auto-resets, including a
    ragged last block (n not a multiple of 32)."""
    from gym_ballenv_b200 import BallVecEnv
    from oracle.gen_golden import CFG_DEFAULT, CFG_DENSE
    cfg = CFG_DEFAULT if cfgname == "default" else CFG_DENSE
    if no_lean:   # the block-of-roles specialisation instead of the thread-per-environment kernel
        monkeypatch.setenv("BALLENV_NO_LEAN", "1")
    fast = BallVecEnv(n, window=w, config=_env_config(cfg), seed=5, max_episode_steps=15)
    assert fast.kernel_variant(1) == ("roles" if no_lean or w == 7 else "lean")
    monkeypatch.setenv("BALLENV_FORCE_GENERIC", "1")
    slow = BallVecEnv(n, window=w, config=_env_config(cfg), seed=5, max_episode_steps=15)
    monkeypatch.delenv("BALLENV_FORCE_GENERIC")
    assert torch.equal(fast.reset(), slow.reset())
    g = torch.Generator().manual_seed(11)
    for t in range(50):
        a = torch.randint(0, 9, (n,), generator=g).cuda()
        of, rf, df, inf = fast.step(a)
        os_, rs, ds, ins = slow.step(a)
        assert torch.equal(of, os_), t
        assert torch.equal(rf, rs), t
        assert torch.equal(df, ds), t
        assert torch.equal(inf["flags"], ins["flags"]), t
    sf, ss = fast.get_state(), slow.get_state()
    for k in sf:
        assert torch.equal(sf[k], ss[k]), k
    stf, sts = fast.stats(), slow.stats()
    for k in stf:   # sums of doubles are accumulated with atomics: order, hence the last bits, may differ
        assert stf[k] == pytest.approx(sts[k], rel=1e-12), k
    fast.close()
    slow.close()


@pytest.mark.gpu
@pytest.mark.parametrize("no_lean", [False, True])
@pytest.mark.parametrize("w,cfgname,n,keep", [(5, "default", 1000, True), (10, "dense", 777, True), (5, "default", 96, False)])
def test_rollout_kernel_matches_per_step_launches(w, cfgname, n, keep, no_lean, monkeypatch):
    """ballenv_step_many as ONE launch (state held on chip for all T steps) == T single-step launches: every
    observation, reward, done, the final state and the statistics, with auto-resets inside the rollout."""
    from gym_ballenv_b200 import BallVecEnv
    from oracle.gen_golden import CFG_DEFAULT, CFG_DENSE
    cfg = CFG_DEFAULT if cfgname == "default" else CFG_DENSE
    T = 64
    if no_lean:
        monkeypatch.setenv("BALLENV_NO_LEAN", "1")
    one = BallVecEnv(n, window=w, config=_env_config(cfg), seed=21, max_episode_steps=17)
    monkeypatch.setenv("BALLENV_NO_ROLLOUT", "1")
    per = BallVecEnv(n, window=w, config=_env_config(cfg), seed=21, max_episode_steps=17)
    monkeypatch.delenv("BALLENV_NO_ROLLOUT")
    assert torch.equal(one.reset(), per.reset())
    g = torch.Generator().manual_seed(5)
    for chunk in range(2):
        a = torch.randint(0, 9, (T, n), generator=g).cuda()
        o1, r1, d1 = one.step_many(a, keep_all_obs=keep)
        l0 = per.launch_count
        o2, r2, d2 = per.step_many(a, keep_all_obs=keep)
        assert per.launch_count - l0 == T
        assert torch.equal(o1, o2)
        assert torch.equal(r1, r2)
        assert torch.equal(d1, d2)
    assert one.launch_count == 1 + 2          # reset + one launch per step_many call
    s1, s2 = one.get_state(), per.get_state()
    for k in s1:
        assert torch.equal(s1[k], s2[k]), k
    st1, st2 = one.stats(), per.stats()
    for k in st1:
        assert st1[k] == pytest.approx(st2[k], rel=1e-12), k
    assert one.error_flags() == 0 and per.error_flags() == 0
    one.close()
    per.close()


@pytest.mark.gpu
@pytest.mark.parametrize("parity", [True, False])
def test_pygame_ruleset_against_the_reference_rollout(parity):
    """createBoard (ballenv_pygame.py:650-706, 460-513) recorded through the shim, raw float actions, auto-reset, the
    whole rollout including the resets.  Positions and done bit-exact; distances / rewards to 1e-12 relative - for
    non-integral coordinates glibc's pow(x, 2) (what math.pow calls) is not guaranteed correctly rounded, the kernel's
    x * x is, so a distance may differ in the last bit ("fp64 parity mode").  The pygame ruleset
    stores fp64 whatever precision is asked for (its rewards are differences of nearby distances of non-integral
    points: fp32 positions cannot hold the 1e-5 relative tolerance), so `parity=False` runs the same arithmetic: the
    rewards are inside north_star's 1e-5 - by eleven orders of magnitude."""
    from gym_ballenv_b200 import BallVecEnv, EnvConfig
    z, meta = load_golden("rollout_pygame")
    n, T, g0 = meta["n_envs"], meta["T"], meta["g0"]
    env = BallVecEnv(n, window=5, config=EnvConfig.pygame_default(static_obstacles=meta["n_static"]), ruleset="pygame",
                     seed=meta["seed"], parity=parity, global_env_offset=g0)
    env.reset()
    st = env.get_state()
    real = np.float64
    assert env.parity and st["agent_x"].dtype == torch.float64
    parity = True
    assert np.array_equal(st["agent_x"].cpu().numpy(), z["init_agent"][:, 0].astype(real))
    assert np.array_equal(st["goal_y"].cpu().numpy(), z["init_goal"][:, 1].astype(real))
    assert np.array_equal(st["static_x"].cpu().numpy().T, z["init_obst"][:, :, 0].astype(real))
    if parity:
        np.testing.assert_allclose(st["dist"].cpu().numpy(), z["init_dist"], rtol=1e-14)
        np.testing.assert_allclose(st["total_distance"].cpu().numpy(), z["init_total_distance"], rtol=1e-14)
    for t in range(T):
        a = torch.from_numpy(z["rec_actions"][t].astype(real)).cuda()
        obs, rew, done, info = env.step(a)
        rew, done = rew.cpu().numpy(), done.cpu().numpy()
        st = env.get_state()
        if parity:
            np.testing.assert_allclose(rew, z["rec_reward"][t], rtol=1e-12, atol=1e-15)
            assert np.array_equal(done, z["rec_done"][t].astype(bool)), t
            assert np.array_equal(st["agent_x"].cpu().numpy(), z["rec_agent"][t, :, 0]), t
            assert np.array_equal(st["agent_y"].cpu().numpy(), z["rec_agent"][t, :, 1]), t
            np.testing.assert_allclose(st["dist"].cpu().numpy(), z["rec_dist"][t], rtol=1e-14)
            assert np.array_equal(st["static_y"].cpu().numpy().T, z["rec_obst"][t, :, :, 1]), t
            nd = ~done
            np.testing.assert_allclose(st["acc_reward"].cpu().numpy()[nd], z["rec_acc"][t][nd], rtol=1e-11, atol=1e-14)
        # north_star: 1e-5 relative (the absolute floor only covers rewards that are zero up to rounding)
        np.testing.assert_allclose(rew, z["rec_reward"][t], rtol=1e-5, atol=1e-12)
    if parity:
        assert env.stats()["episodes"] == meta["episodes"]
    assert env.error_flags() == 0
    env.close()


@pytest.mark.gpu
def test_reset_fixed_against_the_reference_and_the_oracle():
    """ballenv_reset_fixed == createBoard.resetFixedstate (ballenv_pygame.py:589-624): the reference-recorded fixture
    (goal (145, 120): agent, state[2], total_distance, zeroed accumulated reward, obstacles kept, the rewards of the
    steps that follow), then a goal in the middle of the world, where the inner redraw-while-closer-than-50 loop fires
    for most environments, against the oracle's restatement; the facade method on top."""
    from gym_ballenv_b200 import BallVecEnv, EnvConfig, createBoard
    from oracle import draws as D
    from oracle.ballenv_oracle import RULESET_PYGAME, OracleConfig, OracleVec
    z, meta = load_golden("reset_fixed_kat")
    n, g0, ks = meta["n_envs"], meta["g0"], meta["n_static"]
    env = BallVecEnv(n, window=5, config=EnvConfig.pygame_default(static_obstacles=ks), ruleset="pygame", seed=meta["seed"],
                     global_env_offset=g0, auto_reset=False, max_episode_steps=0)
    env.reset()
    for k in range(meta["steps"]):
        env.step(torch.from_numpy(z["pre_actions"][k]).cuda())
    for r in range(meta["rounds"]):
        obs = env.reset_fixed(tuple(meta["goal"]))
        st = env.get_state()
        assert np.array_equal(st["agent_x"].cpu().numpy(), z["rec_agent"][r, :, 0])
        assert np.array_equal(st["agent_y"].cpu().numpy(), z["rec_agent"][r, :, 1])
        assert np.array_equal(st["goal_x"].cpu().numpy(), z["rec_goal"][r, :, 0])
        np.testing.assert_allclose(st["dist"].cpu().numpy(), z["rec_dist"][r], rtol=1e-14)
        np.testing.assert_allclose(st["total_distance"].cpu().numpy(), z["rec_total_distance"][r], rtol=1e-14)
        assert np.array_equal(st["acc_reward"].cpu().numpy(), z["rec_acc"][r]) and int(st["ep_len"].max()) == 0
        assert np.array_equal(st["static_x"].cpu().numpy().T, z["rec_obst"][r, :, :, 0])
        assert torch.equal(obs, env.observe()) and torch.all(obs[:, :4].sum(1) == 1)
        for k in range(meta["steps"]):
            _, rew, done, _ = env.step(torch.from_numpy(z["rec_actions"][r, k]).cuda())
            np.testing.assert_allclose(rew.cpu().numpy(), z["rec_reward"][r, k], rtol=1e-10, atol=1e-15)
            assert np.array_equal(done.cpu().numpy(), z["rec_done"][r, k].astype(bool)), (r, k)
    # a goal inside the world: most first draws are closer than 50 -> the inner loop; masked: only the even environments
    cfg = OracleConfig(ruleset=RULESET_PYGAME, window=5, n_static=ks, n_dynamic=0, speeds=(), goals=(), max_episode_steps=0,
                       auto_reset=False)
    vec = OracleVec(cfg, D.PhiloxDraws(5), n, 0)
    env2 = BallVecEnv(n, window=5, config=EnvConfig.pygame_default(static_obstacles=ks), ruleset="pygame", seed=5,
                      auto_reset=False, max_episode_steps=0)
    env2.reset()
    vec.reset()
    before = env2.get_state()
    mask = torch.arange(n) % 2 == 0
    env2.reset_fixed((60.0, 40.0), mask=mask)
    st = env2.get_state()
    redrawn = 0
    for i, e in enumerate(vec.envs):
        if i % 2 == 0:
            e.reset_fixed((60.0, 40.0))
            assert (st["agent_x"][i].item(), st["agent_y"][i].item()) == tuple(e.agent), i
            assert st["dist"][i].item() == pytest.approx(e.dist, rel=1e-14)
            assert st["total_distance"][i].item() == pytest.approx(e.total_distance, rel=1e-14)
            assert st["episode"][i].item() == e.episode == 1
            redrawn += e.dist != e.total_distance
        else:
            assert st["agent_x"][i].item() == before["agent_x"][i].item() and st["episode"][i].item() == 0
    assert redrawn > n // 8
    assert env.error_flags() == 0 and env2.error_flags() == 0
    env.close()
    env2.close()
    board = createBoard(static_obstacles=ks, seed=meta["seed"] + 1)
    with pytest.raises(RuntimeError):
        board.resetFixedstate()
    board.reset()
    s = board.resetFixedstate()
    assert s[1] == (145, 120) and board.total_reward_accumulated == 0 and len(s) == 3 + ks
    assert board.total_distance == pytest.approx(board.calculate_distance(s[0], s[1]), rel=1e-14)
    board.close()


@pytest.mark.gpu
def test_rollout_buffer_with_unaligned_steps():
    """[T][n][row] rollout rows of step t start at t * n * row elements: with n = 777 and W = 5 (row = 29) that is
    not a multiple of 4 elements, so the 128-bit store path must fall back to element stores for those steps."""
    from gym_ballenv_b200 import BallVecEnv
    n, T = 777, 9
    a = torch.randint(0, 9, (T, n), generator=torch.Generator().manual_seed(2)).cuda()
    envs = [BallVecEnv(n, window=5, seed=4, max_episode_steps=5) for _ in range(2)]
    for e in envs:
        e.reset()
    o1, r1, d1 = envs[0].step_many(a, keep_all_obs=True)
    ref = []
    envs[1].close()
    e2 = BallVecEnv(n, window=5, seed=4, max_episode_steps=5)
    e2.reset()
    for t in range(T):
        o, r, d, _ = e2.step(a[t])
        ref.append((o.clone(), r.clone(), d.clone()))
    assert torch.equal(o1, torch.stack([x[0] for x in ref]))
    assert torch.equal(r1, torch.stack([x[1] for x in ref]))
    assert torch.equal(d1, torch.stack([x[2] for x in ref]))
    assert envs[0].error_flags() == 0


@pytest.mark.gpu
@pytest.mark.parametrize("parity", [True, False])
def test_features20_kernel_against_the_reference_vectors(parity):
    """ballenv_observe_features vs the 20 floats the reference's featureExtractor helpers produce for 160 injected
    states (tests/golden/features_kat.npz).  Counts and one-hots exact; the social-force sums (exp, cos, acos) to
    1e-12 relative in fp64 parity mode and 2e-5 in fp32 (transcendental functions, fp32 positions)."""
    from gym_ballenv_b200 import BallVecEnv, EnvConfig
    z, meta = load_golden("features_kat")
    n, K = meta["n"], meta["K"]
    for rad in (5, 10):
        sel = np.where(z["agent_rad"] == rad)[0]
        env = BallVecEnv(len(sel), window=5, config=EnvConfig.pygame_default(static_obstacles=K, agent_radius=rad),
                         ruleset="pygame", parity=parity, auto_reset=False)
        env.reset()
        env.set_state(agent_x=z["agent"][sel, 0], agent_y=z["agent"][sel, 1], goal_x=z["goal"][sel, 0],
                      goal_y=z["goal"][sel, 1], static_x=z["obst"][sel, :, 0].T, static_y=z["obst"][sel, :, 1].T)
        got = env.sensor_readings().cpu().numpy().astype(np.float64)
        want = z["features"][sel]
        if parity:
            np.testing.assert_allclose(got, want.astype(np.float32), rtol=1e-6, atol=0)     # output rows are fp32
            assert np.array_equal(got[:, :17], want[:, :17])
        else:
            # fp32 positions: a state exactly on a bin boundary may fall on the other side; none of the fixtures is
            assert np.array_equal(got[:, :17], want[:, :17])
            np.testing.assert_allclose(got[:, 17:], want[:, 17:], rtol=2e-5, atol=1e-5)
        env.close()


@pytest.mark.gpu
def test_createboard_facade_against_the_reference_rollout():
    """gym_ballenv_b200.createBoard (reset / step / state list / sensor_readings) replays one environment of the
    reference createBoard rollout (the first one whose episode ends): same draws (seed, global id), same actions."""
    import gym_ballenv_b200 as gb
    from oracle.ballenv_oracle import features20
    z, meta = load_golden("rollout_pygame")
    i = int(np.argmax(z["rec_done"].any(0)))
    env = gb.createBoard(display=False, static_obstacles=meta["n_static"], seed=meta["seed"])
    env._vec.close()
    env._vec = gb.BallVecEnv(1, window=5, config=gb.EnvConfig.pygame_default(static_obstacles=meta["n_static"]),
                             ruleset="pygame", seed=meta["seed"], parity=True, auto_reset=False, max_episode_steps=0,
                             global_env_offset=meta["g0"] + i)
    state = env.reset()
    assert state[0] == tuple(z["init_agent"][i]) and state[1] == tuple(z["init_goal"][i])
    assert [s[:2] for s in state[3:]] == [tuple(int(v) for v in o) for o in z["init_obst"][i]] and state[3][2] == 20
    assert env.total_distance == pytest.approx(z["init_total_distance"][i], rel=1e-14)
    for t in range(meta["T"]):
        state, reward, done, info = env.step(tuple(z["rec_actions"][t, i]))
        assert reward == pytest.approx(z["rec_reward"][t, i], rel=1e-12, abs=1e-15) and done == bool(z["rec_done"][t, i])
        sr = env.sensor_readings
        assert tuple(sr.shape) == (1, 20) and sr.is_cuda
        want = features20(state[0], state[1], [s[:2] for s in state[3:]], agent_rad=10)
        np.testing.assert_allclose(sr.cpu().numpy()[0], np.array(want, dtype=np.float32), rtol=1e-6)
        if done:
            break
        assert state[0] == tuple(z["rec_agent"][t, i])
    assert done and env.actionArray[1].tolist() == [1, 0]
    env.close()


@pytest.mark.gpu
@pytest.mark.parametrize("obs_dtype", [torch.float32, torch.uint8])
def test_step_host_equals_step(obs_dtype):
    """ballenv_step_host (host buffers, copies inside the call - the path bench.py's e2e goes through) returns what
    ballenv_step returns, for fp32 and uint8 observation rows."""
    from gym_ballenv_b200 import BallVecEnv
    n = 333
    dev_env = BallVecEnv(n, window=5, seed=8, max_episode_steps=9, obs_dtype=obs_dtype)
    host_env = BallVecEnv(n, window=5, seed=8, max_episode_steps=9, obs_dtype=obs_dtype)
    dev_env.reset()
    host_env.reset()
    obs = torch.empty((n, 29), dtype=obs_dtype).pin_memory()
    rew = torch.empty(n, dtype=torch.float32).pin_memory()
    done = torch.empty(n, dtype=torch.uint8).pin_memory()
    g = torch.Generator().manual_seed(6)
    for t in range(25):
        a = torch.randint(0, 9, (n,), generator=g)
        o, r, d, _ = dev_env.step(a.cuda())
        host_env.step_host(a.pin_memory(), obs, rew, done)
        assert torch.equal(o.cpu(), obs) and torch.equal(r.cpu(), rew) and torch.equal(d.cpu(), done.bool()), t
    assert host_env.error_flags() == 0


@pytest.mark.gpu
@pytest.mark.parametrize("obs_dtype,parity,w", [(torch.float32, False, 10), (torch.uint8, False, 5), ("bits", False, 10),
                                                (torch.float32, True, 5)])
def test_step_many_host_equals_step_many(obs_dtype, parity, w):
    """ballenv_step_many_host (host buffers; chunks of steps, one rollout launch per chunk, copies overlapped on two
    internal streams and two staging sets) returns what the same steps return on the device - several chunks, so that
    a staging set is reused; production (lean rollout kernel) and fp64 parity (one generic launch per step) handles."""
    from gym_ballenv_b200 import BallVecEnv
    n, T = 333, 150
    kw = dict(window=w, seed=8, max_episode_steps=11, obs_dtype=obs_dtype, parity=parity)
    dev_env, host_env = BallVecEnv(n, **kw), BallVecEnv(n, **kw)
    dev_env.reset()
    host_env.reset()
    g = torch.Generator().manual_seed(6)
    for rep in range(2):
        a = torch.randint(0, 9, (T, n), generator=g)
        o, r, d = dev_env.step_many(a.cuda(), keep_all_obs=True)
        obs = torch.empty(tuple(o.shape), dtype=o.dtype).pin_memory()
        rew = torch.empty((T, n), dtype=r.dtype).pin_memory()
        done = torch.empty((T, n), dtype=torch.uint8).pin_memory()
        host_env.step_many_host(a.pin_memory(), obs, rew, done)
        assert torch.equal(o.cpu(), obs) and torch.equal(r.cpu(), rew) and torch.equal(d.cpu(), done.bool()), rep
    sa, sb = dev_env.get_state(), host_env.get_state()
    for k in sa:
        assert torch.equal(sa[k], sb[k]), k
    assert host_env.error_flags() == 0
    dev_env.close()
    host_env.close()


@pytest.mark.gpu
def test_state_snapshot_resumes_identically():
    """Checkpoint / resume of the environment state: get_state() of a running env loaded into a
    fresh env of the same seed with set_state() continues bit-identically, resets included."""
    from gym_ballenv_b200 import BallVecEnv
    n = 500
    a = torch.randint(0, 9, (45, n), generator=torch.Generator().manual_seed(3)).cuda()
    env = BallVecEnv(n, window=5, seed=12, max_episode_steps=13)
    env.reset()
    env.step_many(a[:15])
    snap = env.get_state()
    o1, r1, d1 = env.step_many(a[15:], keep_all_obs=True)
    env2 = BallVecEnv(n, window=5, seed=12, max_episode_steps=13)
    env2.reset()
    env2.set_state(**snap)
    o2, r2, d2 = env2.step_many(a[15:], keep_all_obs=True)
    assert torch.equal(o1, o2) and torch.equal(r1, r2) and torch.equal(d1, d2)
    s1, s2 = env.get_state(), env2.get_state()
    for k in s1:
        assert torch.equal(s1[k], s2[k]), k


@pytest.mark.gpu
@pytest.mark.parametrize("w", [1, 2, 3, 16, 31, 32])
def test_extreme_window_sizes_against_the_oracle(w):
    """WINDOW is a free parameter in the reference (5, 10, 21, 50 appear in its scripts); the kernel supports 1..32.
    Small rollouts with auto-reset against the Python oracle, per-step launches and the rollout kernel."""
    from gym_ballenv_b200 import BallVecEnv
    from oracle import draws as D
    from oracle.ballenv_oracle import OracleConfig, OracleVec
    n, T, seed = 40, 12, 5
    env = BallVecEnv(n, window=w, seed=seed, max_episode_steps=7)
    vec = OracleVec(OracleConfig(window=w, max_episode_steps=7), D.PhiloxDraws(seed), n)
    obs = env.reset()
    vec.reset()
    assert np.array_equal(obs.cpu().numpy(), np.array(vec.observe(), dtype=np.float32))
    g = torch.Generator().manual_seed(w)
    a = torch.randint(0, 9, (2 * T, n), generator=g)
    for t in range(T):
        obs, rew, done, _ = env.step(a[t].cuda())
        r, d, f = vec.step(a[t].tolist())
        assert np.array_equal(obs.cpu().numpy(), np.array(vec.observe(), dtype=np.float32)), t
        assert np.array_equal(done.cpu().numpy(), np.array(d)), t
    obs, rew, done = env.step_many(a[T:].cuda(), keep_all_obs=True)
    for t in range(T):
        r, d, f = vec.step(a[T + t].tolist())
        assert np.array_equal(obs[t].cpu().numpy(), np.array(vec.observe(), dtype=np.float32)), t
        assert np.array_equal(done[t].cpu().numpy(), np.array(d)), t
    assert env.error_flags() == 0
    with pytest.raises(Exception):
        BallVecEnv(4, window=33)


@pytest.mark.gpu
def test_integer_sqrt_shortcut_is_exact():
    """The step kernel takes sqrt_int22 for the distance to the goal when every coordinate is integral: it must
    return the bits of sqrt() for every argument it can see (squared distances of a world of at most 1024 x 1024)."""
    import ctypes as C
    from gym_ballenv_b200 import _lib as L
    bad = C.c_int64(-1)
    rc = L.LIB.ballenv_selftest(0, 1 << 22, 0, C.byref(bad))
    assert rc == 0, L.LIB.ballenv_last_error()
    assert bad.value == 0, "%d of 2^22 arguments differ from sqrt()" % bad.value


@pytest.mark.gpu
def test_inline_reward_division_is_exact():
    """The lean kernels divide the progress reward with the fp64 division's fast path written out (no out-of-line
    slow path in the step loop): bit-identical to `/` on 2^28 operand pairs shaped like the reward's."""
    import ctypes as C
    from gym_ballenv_b200 import _lib as L
    bad = C.c_int64(-1)
    rc = L.LIB.ballenv_selftest(1, 1 << 28, 0, C.byref(bad))
    assert rc == 0, L.LIB.ballenv_last_error()
    assert bad.value == 0, "%d of 2^28 quotients differ from /" % bad.value


@pytest.mark.gpu
@pytest.mark.parametrize("parity", [False, True])
def test_near_list_overflow(parity):
    """Every obstacle of every environment inside the agent's window: 32 x 32 = 1024 near obstacles per block, more
    than the shared-memory near list holds, so the in-thread raster of the overflow path produces part of the rows."""
    from gym_ballenv_b200 import BallVecEnv, EnvConfig
    from oracle.ballenv_oracle import window_obs
    n, k, w = 64, 32, 10
    rng = np.random.RandomState(3)
    agent = rng.randint(60, 440, size=(n, 2)).astype(np.float64)
    goal = rng.randint(0, 500, size=(n, 2)).astype(np.float64)
    obst = agent[:, None, :] + rng.randint(-30, 31, size=(n, k, 2))
    cfg = EnvConfig(static_obstacles=k, dynamic_obstacles=0, obstacle_speed=(), obs_goal_position=())
    env = BallVecEnv(n, window=w, config=cfg, parity=parity, auto_reset=False, max_episode_steps=0)
    env.set_state(agent_x=agent[:, 0], agent_y=agent[:, 1], goal_x=goal[:, 0], goal_y=goal[:, 1],
                  static_x=obst[:, :, 0].T, static_y=obst[:, :, 1].T)
    obs = env.observe().cpu().numpy()
    for e in range(n):
        exp = window_obs(tuple(agent[e]), tuple(goal[e]), [tuple(o) for o in obst[e]], w)
        assert np.array_equal(obs[e], np.asarray(exp, dtype=obs.dtype)), e
    assert env.error_flags() == 0
    env.close()


@pytest.mark.gpu
@pytest.mark.parametrize("parity", [True, False])
def test_block_counts_kernel_against_the_reference_vectors(parity):
    """ballenv_observe_blocks vs prep_state2 of examples/ball_env_reinforce.py for 200 injected states
    (tests/golden/blocks_kat.npz): counts and one-hots, bit-exact (the non-integral fixtures are exact in fp32)."""
    from gym_ballenv_b200 import BallVecEnv, EnvConfig
    z, meta = load_golden("blocks_kat")
    n, K = meta["n"], meta["K"]
    half = K // 2   # first half static, second half dynamic: the count runs over both lists in order
    cfg = EnvConfig(static_obstacles=half, dynamic_obstacles=K - half, obstacle_speed=(1,) * (K - half),
                    obs_goal_position=tuple("%d,%d" % (10 + 7 * j, 20 + 3 * j) for j in range(K - half)))
    env = BallVecEnv(n, window=5, config=cfg, parity=parity, auto_reset=False, max_episode_steps=0)
    env.reset()
    agent, goal, obst = z["agent"], z["goal"], z["obst"]
    if not parity:   # fp32 state: keep the fixtures whose coordinates survive the cast
        ok = np.all(agent == agent.astype(np.float32), 1) & np.all(goal == goal.astype(np.float32), 1) & \
            np.all(obst == obst.astype(np.float32), (1, 2))
        assert ok.sum() >= 150
    else:
        ok = np.ones(n, bool)
    env.set_state(agent_x=agent[:, 0], agent_y=agent[:, 1], goal_x=goal[:, 0], goal_y=goal[:, 1],
                  static_x=obst[:, :half, 0].T, static_y=obst[:, :half, 1].T,
                  dynamic_x=obst[:, half:, 0].T, dynamic_y=obst[:, half:, 1].T)
    got = env.block_counts().cpu().numpy().astype(np.float64)
    assert np.array_equal(got[ok], z["blocks"][ok])
    env.close()


@pytest.mark.gpu
def test_legacy_block_observation_closed_loop_against_the_oracle():
    """rollout_blocks (the REINFORCE scripts' loop on the 29-float observation): a greedy policy drives 256 GPU
    environments for 40 steps; the oracle replays the same actions, and after every step the kernel's block counts
    equal blocks29 of the oracle's state for every environment."""
    from gym_ballenv_b200 import BallVecEnv
    from gym_ballenv_b200.legacy import BlockPolicy
    from oracle import draws as D
    from oracle.ballenv_oracle import OracleVec, blocks29
    from oracle.gen_golden import CFG_DEFAULT
    n, seed = 256, 9
    torch.manual_seed(1)
    policy = BlockPolicy().cuda()
    env = BallVecEnv(n, window=5, config=_env_config(CFG_DEFAULT), seed=seed, parity=True, max_episode_steps=30)
    env.reset()
    orc = OracleVec(oracle_config(CFG_DEFAULT, 5, 30), D.PhiloxDraws(seed), n)
    orc.reset()
    for t in range(40):
        obs = env.block_counts()
        want = np.array([blocks29(st[0], st[1], st[3:]) for st in (e.state() for e in orc.envs)])
        assert np.array_equal(obs.cpu().numpy().astype(np.float64), want), t
        with torch.no_grad():
            action = policy(obs).argmax(-1)
        env.step(action)
        orc.step(action.cpu().numpy())
    env.close()
