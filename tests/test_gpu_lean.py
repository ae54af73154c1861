"""The thread-per-environment kernels (gym_ballenv_b200/csrc/ballenv_lean.cuh) against the block-of-roles
specialisation, the generic kernel and the oracle: same draws, same arithmetic, bit-identical results - on generated
rollouts with auto-resets and on injected states that reach the kernel's rare paths (near-list overflow, counters out
of lockstep, duplicate-free goal changes, non-integral coordinates, ragged last warps)."""
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


def _cfg(name):
    from oracle.gen_golden import CFG_DEFAULT, CFG_DENSE
    return CFG_DEFAULT if name == "default" else CFG_DENSE


def _pair(monkeypatch, n, w, cfgname, other="roles", lanes=0, **kw):
    """(lean env, comparison env) with identical seeds; `other` = "roles" | "generic"; lanes = lanes per environment of
    the lean kernel (0: the library's choice for the configuration)."""
    from gym_ballenv_b200 import BallVecEnv
    cfg = _env_config(_cfg(cfgname))
    if lanes:
        monkeypatch.setenv("BALLENV_LEAN_G", str(lanes))
    lean = BallVecEnv(n, window=w, config=cfg, **kw)
    monkeypatch.delenv("BALLENV_LEAN_G", raising=False)
    monkeypatch.setenv("BALLENV_NO_LEAN" if other == "roles" else "BALLENV_FORCE_GENERIC", "1")
    ref = BallVecEnv(n, window=w, config=cfg, **kw)
    monkeypatch.delenv("BALLENV_NO_LEAN" if other == "roles" else "BALLENV_FORCE_GENERIC")
    assert lean.kernel_variant(1) == "lean" and lean.kernel_variant(8) == "lean"
    assert ref.kernel_variant(1) == other
    return lean, ref


def _same_state(a, b):
    sa, sb = a.get_state(), b.get_state()
    for k in sa:
        assert torch.equal(sa[k], sb[k]), k
    sta, stb = a.stats(), b.stats()
    for k in sta:   # sums of doubles are accumulated with atomics: order, hence the last bits, may differ
        assert sta[k] == pytest.approx(stb[k], rel=1e-12), k
    assert a.error_flags() == 0 and b.error_flags() == 0


@pytest.mark.parametrize("lanes", [1, 2])
@pytest.mark.parametrize("w,cfgname,n,other", [(5, "default", 1000, "roles"), (10, "dense", 777, "roles"),
                                               (10, "default", 33, "generic"), (5, "dense", 65, "generic"),
                                               (10, "dense", 4096, "generic")])
def test_lean_single_step_matches_the_other_kernels(w, cfgname, n, other, lanes, monkeypatch):
    lean, ref = _pair(monkeypatch, n, w, cfgname, other, lanes, seed=5, max_episode_steps=15)
    assert torch.equal(lean.reset(), ref.reset())
    g = torch.Generator().manual_seed(11)
    for t in range(60):
        a = torch.randint(0, 9, (n,), generator=g).cuda()
        ol, rl, dl, il = lean.step(a)
        orf, rr, dr, ir = ref.step(a)
        assert torch.equal(ol, orf), t
        assert torch.equal(rl, rr), t
        assert torch.equal(dl, dr), t
        assert torch.equal(il["flags"], ir["flags"]), t
    _same_state(lean, ref)
    assert lean.stats()["episodes"] > n
    lean.close()
    ref.close()


@pytest.mark.parametrize("lanes", [1, 2])
@pytest.mark.parametrize("w,cfgname,n,keep", [(5, "default", 1000, True), (10, "dense", 777, True), (10, "dense", 96, False),
                                              (5, "dense", 31, True)])
def test_lean_rollout_matches_per_step_launches_of_the_other_kernel(w, cfgname, n, keep, lanes, monkeypatch):
    """ballenv_step_many through the lean rollout kernel (state in registers for all T steps) == T single-step
    launches of the block-of-roles kernel, with auto-resets inside the rollout and action dtypes int64 / int32 / uint8."""
    from gym_ballenv_b200 import BallVecEnv
    cfg = _env_config(_cfg(cfgname))
    T = 64
    monkeypatch.setenv("BALLENV_LEAN_G", str(lanes))
    one = BallVecEnv(n, window=w, config=cfg, seed=21, max_episode_steps=17)
    monkeypatch.delenv("BALLENV_LEAN_G")
    monkeypatch.setenv("BALLENV_NO_LEAN", "1")
    monkeypatch.setenv("BALLENV_NO_ROLLOUT", "1")
    per = BallVecEnv(n, window=w, config=cfg, seed=21, max_episode_steps=17)
    monkeypatch.delenv("BALLENV_NO_LEAN")
    monkeypatch.delenv("BALLENV_NO_ROLLOUT")
    assert one.kernel_variant(T) == "lean" and per.kernel_variant(T) == "roles"
    assert torch.equal(one.reset(), per.reset())
    g = torch.Generator().manual_seed(5)
    for chunk, dt in enumerate((torch.int64, torch.int32, torch.uint8)):
        a = torch.randint(0, 9, (T, n), generator=g).to(dt).cuda()
        o1, r1, d1 = one.step_many(a, keep_all_obs=keep)
        o2, r2, d2 = per.step_many(a, keep_all_obs=keep)
        assert torch.equal(o1, o2), chunk
        assert torch.equal(r1, r2), chunk
        assert torch.equal(d1, d2), chunk
    assert one.launch_count == 1 + 3
    _same_state(one, per)
    one.close()
    per.close()


def test_lean_against_the_c_oracle_with_resets():
    """4096 envs x 120 steps, W=10 dense, TimeLimit 13: lean single-step launches vs oracle/ballenv_oracle.c, every
    observation / done / flag byte, then the whole state."""
    from gym_ballenv_b200 import BallVecEnv
    from oracle.c_oracle import COracleVec
    from oracle.gen_golden import CFG_DENSE
    n, T, seed = 4096, 120, 99
    env = BallVecEnv(n, window=10, config=_env_config(CFG_DENSE), seed=seed, max_episode_steps=13)
    assert env.kernel_variant(1) == "lean"
    c = COracleVec(oracle_config(CFG_DENSE, 10, 13), seed, n)
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
    st, ref = env.get_state(), c.state()
    assert np.array_equal(st["dist"].cpu().numpy(), ref["dist"])
    assert np.array_equal(st["dynamic_x"].cpu().numpy().T, ref["obstacles"][:, 8:, 0].astype(np.float32))
    assert np.array_equal(st["dynamic_goal"].cpu().numpy().T, ref["dyn_goal"])
    assert np.array_equal(st["dynamic_counter"].cpu().numpy().T, ref["dyn_counter"])
    assert env.error_flags() == 0
    env.close()


def _inject(envs, **fields):
    for e in envs:
        e.set_state(**fields)


@pytest.mark.parametrize("lanes", [1, 2])
@pytest.mark.parametrize("w,cfgname", [(10, "dense"), (5, "default")])
def test_lean_near_list_overflow_and_hits(w, cfgname, lanes, monkeypatch):
    """Every obstacle of every environment piled onto the agent's window: more near obstacles than the per-lane list
    holds (rescan path), hits on the first step, resets right after."""
    n = 200
    lean, ref = _pair(monkeypatch, n, w, cfgname, "generic", lanes, seed=3, max_episode_steps=40)
    lean.reset()
    ref.reset()
    rng = np.random.RandomState(3)
    ks, kd = lean.config.static_obstacles, lean.config.dynamic_obstacles
    agent = rng.randint(60, 440, size=(n, 2)).astype(np.float32)
    sx = agent[:, 0][None, :] + rng.randint(-33, 34, size=(ks, n))
    sy = agent[:, 1][None, :] + rng.randint(-33, 34, size=(ks, n))
    dx = agent[:, 0][None, :] + rng.randint(-33, 34, size=(kd, n))
    dy = agent[:, 1][None, :] + rng.randint(-33, 34, size=(kd, n))
    _inject((lean, ref), agent_x=agent[:, 0], agent_y=agent[:, 1], static_x=sx, static_y=sy, dynamic_x=dx, dynamic_y=dy)
    g = torch.Generator().manual_seed(2)
    for t in range(6):
        a = torch.randint(0, 9, (n,), generator=g).cuda()
        ol, rl, dl, il = lean.step(a)
        orf, rr, dr, ir = ref.step(a)
        assert torch.equal(ol, orf), t
        assert torch.equal(rl, rr), t
        assert torch.equal(dl, dr), t
        assert torch.equal(il["flags"], ir["flags"]), t
    _same_state(lean, ref)
    lean.close()
    ref.close()


@pytest.mark.parametrize("lanes", [1, 2])
def test_lean_counters_out_of_lockstep_and_stale_counters(lanes, monkeypatch):
    """Injected change counters that differ inside a quad (the per-obstacle path), sit at the change step, or lie
    beyond it (treated like the change step, as the reference's `<` test does)."""
    n, w = 96, 10
    lean, ref = _pair(monkeypatch, n, w, "dense", "generic", lanes, seed=8, max_episode_steps=0, auto_reset=False)
    lean.reset()
    ref.reset()
    rng = np.random.RandomState(1)
    kd = lean.config.dynamic_obstacles
    cnt = rng.randint(0, 52, size=(kd, n))
    cnt[:, :8] = 50                     # whole environments at the change step
    cnt[::3, 8:16] = 50                 # mixed quads
    cnt[1, 16:24] = 300                 # stale counter beyond the change step
    _inject((lean, ref), dynamic_counter=cnt)
    g = torch.Generator().manual_seed(4)
    for t in range(110):
        a = torch.randint(0, 9, (n,), generator=g).cuda()
        ol, rl, dl, il = lean.step(a)
        orf, rr, dr, ir = ref.step(a)
        assert torch.equal(ol, orf), t
        assert torch.equal(rl, rr), t
    sl, sr = lean.get_state(), ref.get_state()
    for k in sl:
        assert torch.equal(sl[k], sr[k]), k
    lean.close()
    ref.close()


@pytest.mark.parametrize("lanes", [1, 2])
def test_lean_non_integral_coordinates_take_the_general_path(lanes, monkeypatch):
    """Fractional agent / obstacle positions (injected; the gym ruleset never produces them): the table raster and the
    integer square root do not apply, the per-cell arithmetic of the generic kernel does - same fp32 results as the
    block-of-roles kernel, which shares that arithmetic."""
    n, w = 128, 10
    lean, ref = _pair(monkeypatch, n, w, "dense", "roles", lanes, seed=12, max_episode_steps=0, auto_reset=False)
    lean.reset()
    ref.reset()
    st = lean.get_state()
    rng = np.random.RandomState(5)
    ax = st["agent_x"].cpu().numpy() + rng.randint(0, 4, size=n) * 0.25
    dx = st["dynamic_x"].cpu().numpy() + rng.randint(0, 4, size=st["dynamic_x"].shape) * 0.25
    sy = st["static_y"].cpu().numpy()
    sy[:, : n // 2] = st["agent_y"].cpu().numpy()[None, : n // 2] + 7.5     # near, fractional
    _inject((lean, ref), agent_x=ax, dynamic_x=dx, static_y=sy)
    g = torch.Generator().manual_seed(6)
    a = torch.randint(0, 9, (30, n), generator=g).cuda()
    o1, r1, d1 = lean.step_many(a, keep_all_obs=True)
    o2, r2, d2 = ref.step_many(a, keep_all_obs=True)
    assert torch.equal(o1, o2)
    assert torch.equal(r1, r2)
    assert torch.equal(d1, d2)
    assert o1[:, :, 4:].sum() > 0
    _same_state(lean, ref)
    lean.close()
    ref.close()


def test_lean_is_not_selected_outside_its_domain():
    from gym_ballenv_b200 import BallVecEnv, EnvConfig
    env = BallVecEnv(64, window=7)                                   # window without an instantiation
    assert env.kernel_variant(1) == "roles"
    env.close()
    env = BallVecEnv(64, window=5, config=EnvConfig(time_step_for_change=300))   # counter does not fit a byte
    assert env.kernel_variant(1) == "roles"
    env.close()
    env = BallVecEnv(64, window=5, parity=True)                      # fp64 parity mode
    assert env.kernel_variant(1) == "generic"
    env.close()
    env = BallVecEnv(64, window=5)
    assert env.kernel_variant(1) == "lean" and env.kernel_variant(1, torch.float32) == "generic"
    env.close()


def test_direct_state_writes_need_state_written(monkeypatch):
    """The lean kernels skip their integrality test while the library knows the state to be integral (full reset, index
    actions only).  A caller that edits the live state views directly announces it with state_written(): the state is
    re-validated on the device and fractional coordinates take the general path again - same results as the
    block-of-roles kernel, which shares that arithmetic."""
    n, w = 256, 10
    lean, ref = _pair(monkeypatch, n, w, "dense", "roles", seed=2, max_episode_steps=0, auto_reset=False)
    lean.reset()
    ref.reset()
    for e in (lean, ref):
        e.state_views["agent_x"][: n // 2] += 0.25        # a direct write into the arena ...
        e.state_views["static_x"][:, ::3] += 0.5
        e.state_written()                                  # ... announced
    g = torch.Generator().manual_seed(9)
    a = torch.randint(0, 9, (40, n), generator=g).cuda()
    o1, r1, d1 = lean.step_many(a, keep_all_obs=True)
    o2, r2, d2 = ref.step_many(a, keep_all_obs=True)
    assert torch.equal(o1, o2) and torch.equal(r1, r2) and torch.equal(d1, d2)
    # integral again (a full reset): the shortcut path, still identical
    assert torch.equal(lean.reset(), ref.reset())
    for t in range(20):
        ol, rl, dl, _ = lean.step(a[t])
        orf, rr, dr, _ = ref.step(a[t])
        assert torch.equal(ol, orf) and torch.equal(rl, rr) and torch.equal(dl, dr), t
    _same_state(lean, ref)
    lean.close()
    ref.close()


@pytest.mark.parametrize("fmt", ["u8", "bits"])
@pytest.mark.parametrize("w,cfgname,n", [(10, "dense", 1000), (5, "default", 777), (5, "dense", 64)])
def test_lean_row_formats(w, cfgname, n, fmt):
    """uint8 and bit-packed observation rows out of the lean kernels (single-step and rollout, ragged last warp) hold
    the same 0 / 1 values as the float32 rows, every step."""
    from gym_ballenv_b200 import BallVecEnv
    cfg = _env_config(_cfg(cfgname))
    dt = {"u8": torch.uint8, "bits": "bits"}[fmt]
    ref = BallVecEnv(n, window=w, config=cfg, seed=5, max_episode_steps=19)
    env = BallVecEnv(n, window=w, config=cfg, seed=5, max_episode_steps=19, obs_dtype=dt)
    assert env.kernel_variant(1) == "lean" and env.kernel_variant(16) == "lean"
    nb = 4 + w * w

    def as01(o):
        o = o.cpu().numpy()
        if fmt == "bits":
            words = o.astype(np.int64) & 0xffffffff
            return ((words[..., None] >> np.arange(32)) & 1).reshape(o.shape[:-1] + (-1,))[..., :nb].astype(np.float32)
        return o.astype(np.float32)

    assert np.array_equal(as01(env.reset()), ref.reset().cpu().numpy())
    g = torch.Generator().manual_seed(0)
    for t in range(25):
        a = torch.randint(0, 9, (n,), generator=g).cuda()
        assert np.array_equal(as01(env.step(a)[0]), ref.step(a)[0].cpu().numpy()), t
    a = torch.randint(0, 9, (33, n), generator=g).cuda()
    o, r, d = env.step_many(a, keep_all_obs=True)
    o2, r2, d2 = ref.step_many(a, keep_all_obs=True)
    assert np.array_equal(as01(o), o2.cpu().numpy()) and torch.equal(r, r2) and torch.equal(d, d2)
    assert env.error_flags() == 0
    env.close()
    ref.close()


# ---------------------------------------------------------------------------------------------------------------------
# run-time obstacle counts (ballenv_lean_kernel<W, -1, -1, G, ..>): the same kernels for every other configuration
def _cfg_counts(ks, kd, change=7):
    from gym_ballenv_b200 import EnvConfig
    goals = ["%d,%d" % (30 + (53 * i) % 440, 25 + (97 * i) % 450) for i in range(max(kd, 2) + 1)]
    return EnvConfig(static_obstacles=ks, dynamic_obstacles=kd, obstacle_speed=[1 + (j % 3) for j in range(kd)],
                     obs_goal_position=goals, time_step_for_change=change, rd_th_obs=55)


@pytest.mark.parametrize("lanes", [1, 2])
@pytest.mark.parametrize("w,ks,kd,n", [(5, 10, 7, 1000), (10, 3, 2, 333), (5, 0, 1, 64), (10, 1, 9, 200), (5, 31, 33, 150),
                                       (10, 20, 12, 97), (5, 4, 4, 2048)])
def test_runtime_count_lean_kernels_match_the_other_kernels(w, ks, kd, n, lanes, monkeypatch):
    """Obstacle counts without a fixed instance: single-step launches and the rollout loop of the run-time-count lean
    kernels against the block-of-roles kernel (single steps), bit for bit, with resets and goal changes every 7 steps."""
    from gym_ballenv_b200 import BallVecEnv
    cfg = _cfg_counts(ks, kd)
    monkeypatch.setenv("BALLENV_LEAN_G", str(lanes))
    lean = BallVecEnv(n, window=w, config=cfg, seed=9, max_episode_steps=19)
    roll = BallVecEnv(n, window=w, config=cfg, seed=9, max_episode_steps=19)
    monkeypatch.delenv("BALLENV_LEAN_G")
    monkeypatch.setenv("BALLENV_NO_LEAN", "1")
    monkeypatch.setenv("BALLENV_NO_ROLLOUT", "1")
    ref = BallVecEnv(n, window=w, config=cfg, seed=9, max_episode_steps=19)
    monkeypatch.delenv("BALLENV_NO_LEAN")
    monkeypatch.delenv("BALLENV_NO_ROLLOUT")
    assert lean.kernel_variant(1) == "lean" and roll.kernel_variant(40) == "lean" and ref.kernel_variant(1) != "lean"
    assert torch.equal(lean.reset(), ref.reset())
    roll.reset()
    g = torch.Generator().manual_seed(3)
    T = 80
    acts = torch.randint(0, 9, (T, n), generator=g).cuda()
    for t in range(T):
        ol, rl, dl, il = lean.step(acts[t])
        orf, rr, dr, ir = ref.step(acts[t])
        assert torch.equal(ol, orf), t
        assert torch.equal(rl, rr), t
        assert torch.equal(dl, dr), t
        assert torch.equal(il["flags"], ir["flags"]), t
    _same_state(lean, ref)
    o1, r1, d1 = roll.step_many(acts, keep_all_obs=False)
    assert torch.equal(o1, orf)
    _same_state(roll, ref)
    assert lean.stats()["episodes"] > n
    for e in (lean, roll, ref):
        e.close()


@pytest.mark.parametrize("w,cfgname", [(5, "default"), (10, "dense")])
def test_runtime_count_kernels_equal_the_fixed_instances(w, cfgname, monkeypatch):
    """BALLENV_LEAN_RT=1 sends a configuration that HAS a fixed instance through the run-time-count kernel: same results
    (rollout with all rows kept, uint8 and bit-packed rows included)."""
    from gym_ballenv_b200 import BallVecEnv
    cfg = _env_config(_cfg(cfgname))
    n, T = 1500, 70
    for dt in (torch.float32, torch.uint8, "bits"):
        fixed = BallVecEnv(n, window=w, config=cfg, seed=4, max_episode_steps=23, obs_dtype=dt)
        monkeypatch.setenv("BALLENV_LEAN_RT", "1")
        rt = BallVecEnv(n, window=w, config=cfg, seed=4, max_episode_steps=23, obs_dtype=dt)
        monkeypatch.delenv("BALLENV_LEAN_RT")
        assert torch.equal(fixed.reset(), rt.reset())
        a = torch.randint(0, 9, (T, n), generator=torch.Generator().manual_seed(8)).cuda()
        o1, r1, d1 = fixed.step_many(a, keep_all_obs=True)
        o2, r2, d2 = rt.step_many(a, keep_all_obs=True)
        assert torch.equal(o1, o2) and torch.equal(r1, r2) and torch.equal(d1, d2)
        _same_state(fixed, rt)
        fixed.close()
        rt.close()


def test_runtime_count_kernels_against_the_c_oracle():
    """(10, 7) obstacles, W = 5, 2048 environments x 100 single steps with TimeLimit 13 against oracle/ballenv_oracle.c."""
    from gym_ballenv_b200 import BallVecEnv
    from oracle.c_oracle import COracleVec
    from oracle.ballenv_oracle import OracleConfig, RULESET_GYM
    cfg = _cfg_counts(10, 7)
    n, T, seed = 2048, 100, 31
    env = BallVecEnv(n, window=5, config=cfg, seed=seed, max_episode_steps=13)
    assert env.kernel_variant(1) == "lean"
    goals = [tuple(int(v) for v in s.split(",")) for s in cfg.obs_goal_position]
    oc = OracleConfig(ruleset=RULESET_GYM, window=5, n_static=10, n_dynamic=7, speeds=list(cfg.obstacle_speed), goals=goals,
                      change_step=cfg.time_step_for_change, rd_th_obs=cfg.rd_th_obs, static_penalty=cfg.static_penalty[1],
                      dynamic_penalty=cfg.dynamic_penalty[1], max_episode_steps=13, auto_reset=True)
    c = COracleVec(oc, seed, n)
    obs = env.reset()
    c.reset()
    g = torch.Generator().manual_seed(2)
    assert np.array_equal(obs.cpu().numpy(), c.observe())
    for t in range(T):
        a = torch.randint(0, 9, (n,), generator=g)
        obs, rew, done, info = env.step(a.cuda())
        r, d, f = c.step(a.numpy())
        assert np.array_equal(obs.cpu().numpy(), c.observe()), t
        assert np.array_equal(done.cpu().numpy(), d), t
        assert np.array_equal(info["flags"].cpu().numpy(), f), t
        np.testing.assert_allclose(rew.cpu().numpy(), r, rtol=1e-5, atol=0)
    st, ref = env.get_state(), c.state()
    assert np.array_equal(st["dist"].cpu().numpy(), ref["dist"])
    assert np.array_equal(st["dynamic_x"].cpu().numpy().T, ref["obstacles"][:, 10:, 0].astype(np.float32))
    assert np.array_equal(st["dynamic_counter"].cpu().numpy().T, ref["dyn_counter"])
    assert env.error_flags() == 0
    env.close()


@pytest.mark.parametrize("lanes", [1, 2])
def test_runtime_count_kernels_with_fractional_speeds(lanes, monkeypatch):
    """Obstacle speeds 0.5 / 1.25 / 2: the coordinates leave the integers, the exact shortcuts (column-mask table, integer
    square root) must switch themselves off - the run-time-count kernels against the block-of-roles kernel, bit for bit."""
    from gym_ballenv_b200 import BallVecEnv, EnvConfig
    goals = ["%d,%d" % (30 + (53 * i) % 440, 25 + (97 * i) % 450) for i in range(7)]
    cfg = EnvConfig(static_obstacles=6, dynamic_obstacles=6, obstacle_speed=[0.5, 1.25, 2, 0.5, 1, 1.25],
                    obs_goal_position=goals, time_step_for_change=11, rd_th_obs=50)
    n, T = 900, 70
    monkeypatch.setenv("BALLENV_LEAN_G", str(lanes))
    lean = BallVecEnv(n, window=10, config=cfg, seed=17, max_episode_steps=29)
    roll = BallVecEnv(n, window=10, config=cfg, seed=17, max_episode_steps=29)
    monkeypatch.delenv("BALLENV_LEAN_G")
    monkeypatch.setenv("BALLENV_NO_LEAN", "1")
    monkeypatch.setenv("BALLENV_NO_ROLLOUT", "1")
    ref = BallVecEnv(n, window=10, config=cfg, seed=17, max_episode_steps=29)
    monkeypatch.delenv("BALLENV_NO_LEAN")
    monkeypatch.delenv("BALLENV_NO_ROLLOUT")
    assert lean.kernel_variant(1) == "lean" and ref.kernel_variant(1) != "lean"
    assert torch.equal(lean.reset(), ref.reset())
    roll.reset()
    acts = torch.randint(0, 9, (T, n), generator=torch.Generator().manual_seed(1)).cuda()
    for t in range(T):
        ol, rl, dl, _ = lean.step(acts[t])
        orf, rr, dr, _ = ref.step(acts[t])
        assert torch.equal(ol, orf) and torch.equal(rl, rr) and torch.equal(dl, dr), t
    o1, r1, d1 = roll.step_many(acts, keep_all_obs=False)
    assert torch.equal(o1, orf)
    _same_state(lean, ref)
    _same_state(roll, ref)
    assert float(lean.get_state()["dynamic_x"].frac().abs().max()) > 0      # the coordinates did leave the integers
    for e in (lean, roll, ref):
        e.close()
