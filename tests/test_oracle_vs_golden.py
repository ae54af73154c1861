"""CPU: the oracle restatement (oracle/ballenv_oracle.py) against the fixtures
recorded from the reference's own code (oracle/gen_golden.py)."""
import numpy as np
import pytest

from oracle import draws as D
from oracle.ballenv_oracle import (AGENT_MOVES, OracleEnv, OracleVec, goal_quadrant, window_rows)
from helpers import canonical_goal_index, load_golden, oracle_config, tapes_from_golden


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32 10 rounds
    assert D.philox4x32_10(0, 0, 0, 0, 0, 0) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    f = 0xffffffff
    assert D.philox4x32_10(f, f, f, f, f, f) == (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)
    args = (0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0)
    assert D.philox4x32_10(*args) == (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)
    assert tuple(int(x) for x in D.philox4x32_10_np(*args)) == (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)


def test_word_for_roundtrip():
    for n in (2, 4, 9, 10, 20, 100, 460, 500):
        for v in range(n):
            assert D.mulhi(D.word_for(v, n), n) == v
    x = 0.7364512345678
    x = int(x * 2 ** 53) / 2 ** 53
    assert D.ranf_from_words(*D.words_for_ranf(x)) == x


def test_window_kat():
    z, meta = load_golden("window_kat")
    for i in range(len(z["agent"])):
        agent, goal = tuple(z["agent"][i]), tuple(z["goal"][i])
        obst = [tuple(p) for p in z["obst"][i][: z["n_obst"][i]]]
        assert goal_quadrant(agent, goal) == z["quadrant"][i]
        for w in (5, 10, 21):
            assert window_rows(agent, obst, w) == list(z["rows%d" % w][i]), (i, w)


def test_row_offset_quirk():
    # rows 0 and 1 sample the same y (examples/ball_cnn_ac3.py:409): an obstacle only reachable
    # from the top row shows up twice, and the intended last row is never sampled.
    rows = window_rows((100, 100), [(100, 73)], 5)      # y offsets -2,-2,-1,0,1 ; 98-73 = 25 -> hit
    assert rows[0] == rows[1] != 0 and rows[2] == 0


def test_edge_cases():
    z, meta = load_golden("edge_gym")
    cfg = oracle_config(meta["cfg"], 5, max_episode_steps=0, auto_reset=False)
    for i, name in enumerate(meta["names"]):
        tape = z["words"][i].reshape(1, 1, cfg.n_dynamic, 2)
        e = OracleEnv(cfg, D.TapeDraws(step_tape=tape, n_static=cfg.n_static, n_dynamic=cfg.n_dynamic))
        e.agent, e.goal = tuple(z["in_agent"][i]), tuple(z["in_goal"][i])
        e.dist, e.total_distance, e.acc = float(z["in_dist"][i]), float(z["in_total"][i]), meta["in_acc"]
        e.obst = [list(p) for p in z["in_obst"][i]]
        e.goal_idx, e.counter = list(z["in_goal_idx"][i]), list(z["in_counter"][i])
        r, d = e.step(tuple(z["action"][i]))
        assert tuple(e.agent) == tuple(z["out_agent"][i]), name
        assert e.dist == z["out_dist"][i], name
        assert np.array_equal(np.array(e.obst, dtype=np.float64), z["out_obst"][i]), name
        assert e.goal_idx == list(z["out_goal_idx"][i]) and e.counter == list(z["out_counter"][i]), name
        assert r == z["out_reward"][i] and bool(d) == bool(z["out_done"][i]), name
        assert e.acc == z["out_acc"][i], name
        f = (1 if e.goal_flag else 0) | (2 if e.hit else 0) | (8 if e.hit_index >= cfg.n_static else 0)
        assert f == z["out_flags"][i], name
        assert goal_quadrant(e.agent, e.goal) == z["out_quadrant"][i], name
        assert e.observe_rows(5) == list(z["out_rows5"][i]), name
        assert e.observe_rows(10) == list(z["out_rows10"][i]), name


def _check_rollout(name):
    z, meta = load_golden(name)
    w0 = meta["windows"][0]
    cfg = oracle_config(meta["cfg"], w0, meta["max_episode_steps"])
    n, T, g0 = meta["n_envs"], meta["T"], meta["g0"]
    canon = canonical_goal_index(meta["cfg"])   # the fixtures record a goal by its first index (identity unless goals repeat)
    if meta["mode"] == "philox":
        src = D.PhiloxDraws(meta["seed"])
    else:
        step, reset = tapes_from_golden(z, meta)
        src = D.TapeDraws(step, reset, cfg.n_static, cfg.n_dynamic, meta["tape_attempts"], g0=g0)
    vec = OracleVec(cfg, src, n, g0)
    vec.reset()
    for i, e in enumerate(vec.envs):
        assert tuple(e.agent) == tuple(z["init_agent"][i]) and tuple(e.goal) == tuple(z["init_goal"][i])
        assert e.dist == z["init_dist"][i] and e.total_distance == z["init_total_distance"][i]
        assert np.array_equal(np.array(e.obst, dtype=np.float64), z["init_obst"][i].astype(np.float64))
        for w in meta["windows"]:
            assert e.observe_rows(w) == list(z["init_rows%d" % w][i])
    for t in range(T):
        rew, done, flags = vec.step(list(z["rec_actions"][t]))
        assert rew == list(z["rec_reward"][t]), t
        assert [int(d) for d in done] == list(z["rec_done"][t]), t
        assert flags == list(z["rec_flags"][t]), t
        for i, e in enumerate(vec.envs):
            assert tuple(e.agent) == tuple(z["rec_agent"][t, i]) and tuple(e.goal) == tuple(z["rec_goal"][t, i])
            assert e.dist == z["rec_dist"][t, i] and e.acc == z["rec_acc"][t, i]
            assert e.total_distance == z["rec_total_distance"][t, i] and e.ep_len == z["rec_ep_len"][t, i]
            assert np.array_equal(np.array(e.obst, dtype=np.float64), z["rec_obst"][t, i].astype(np.float64)), (t, i)
            assert [int(canon[k]) for k in e.goal_idx] == list(z["rec_dyn_goal"][t, i])
            assert e.counter == list(z["rec_dyn_counter"][t, i])
            assert goal_quadrant(e.agent, e.goal) == z["rec_quadrant"][t, i]
            for w in meta["windows"]:
                assert e.observe_rows(w) == list(z["rec_rows%d" % w][t, i]), (t, i, w)
    for k, v in meta["stats"].items():
        assert vec.stats[k] == pytest.approx(v, rel=1e-12), k
    return meta


@pytest.mark.parametrize("name", ["rollout_philox_default", "rollout_philox_busy", "rollout_philox_dense",
                                  "rollout_mt_default", "rollout_philox_dupgoals", "rollout_checkpoint"])
def test_rollout(name):
    meta = _check_rollout(name)
    assert meta["stats"]["episodes"] > 0


def test_rollout_pygame_ruleset():
    """createBoard (ballenv_pygame.py) stepped through the shim with raw float actions: the oracle's pygame
    ruleset reproduces every position, distance, reward, done and accumulated reward exactly (same fp64 ops)."""
    from oracle.ballenv_oracle import RULESET_PYGAME, OracleConfig
    z, meta = load_golden("rollout_pygame")
    n, T, g0 = meta["n_envs"], meta["T"], meta["g0"]
    cfg = OracleConfig(ruleset=RULESET_PYGAME, window=5, n_static=meta["n_static"], n_dynamic=0, speeds=(), goals=(),
                       max_episode_steps=0, auto_reset=True, agent_radius=meta["agent_radius"],
                       static_obstacle_radius=meta["static_obstacle_radius"])
    vec = OracleVec(cfg, D.PhiloxDraws(meta["seed"]), n, g0)
    vec.reset()
    for i, e in enumerate(vec.envs):
        assert tuple(e.agent) == tuple(z["init_agent"][i]) and tuple(e.goal) == tuple(z["init_goal"][i])
        assert e.dist == z["init_dist"][i] and e.total_distance == z["init_total_distance"][i]
        assert np.array_equal(np.array(e.obst, dtype=np.float64), z["init_obst"][i])
    for t in range(T):
        rew, done, _ = vec.step([tuple(a) for a in z["rec_actions"][t]])
        assert [float(r) for r in rew] == list(z["rec_reward"][t]), t
        assert [int(d) for d in done] == list(z["rec_done"][t]), t
        for i, e in enumerate(vec.envs):
            assert tuple(e.agent) == tuple(z["rec_agent"][t, i]) and tuple(e.goal) == tuple(z["rec_goal"][t, i]), (t, i)
            assert e.dist == z["rec_dist"][t, i] and e.total_distance == z["rec_total_distance"][t, i]
            assert np.array_equal(np.array(e.obst, dtype=np.float64), z["rec_obst"][t, i])
            if not done[i]:
                assert e.acc == z["rec_acc"][t, i]
    assert vec.stats["episodes"] == meta["episodes"] > 0


def test_reset_fixed_against_the_reference():
    """createBoard.resetFixedstate (ballenv_pygame.py:589-624) run on the reference through the shim
    (tests/golden/reset_fixed_kat.npz): goal (145, 120), obstacles kept, the agent redrawn until it is clear of them
    (up to 10 draws in the fixture); the oracle's restatement reproduces agent, state[2], total_distance, the zeroed
    accumulated reward and the rewards of the steps that follow, exactly."""
    from oracle.ballenv_oracle import RULESET_PYGAME, OracleConfig
    z, meta = load_golden("reset_fixed_kat")
    n, g0 = meta["n_envs"], meta["g0"]
    cfg = OracleConfig(ruleset=RULESET_PYGAME, window=5, n_static=meta["n_static"], n_dynamic=0, speeds=(), goals=(),
                       max_episode_steps=0, auto_reset=False)
    vec = OracleVec(cfg, D.PhiloxDraws(meta["seed"]), n, g0)
    vec.reset()
    for k in range(meta["steps"]):
        vec.step([tuple(a) for a in z["pre_actions"][k]])
    assert z["rec_draws"].max() >= 5            # the redraw-while-touching loop is exercised
    for r in range(meta["rounds"]):
        vec.reset_fixed(tuple(meta["goal"]))
        for i, e in enumerate(vec.envs):
            assert tuple(e.agent) == tuple(z["rec_agent"][r, i]) and tuple(e.goal) == tuple(z["rec_goal"][r, i]), (r, i)
            assert e.dist == z["rec_dist"][r, i] and e.total_distance == z["rec_total_distance"][r, i]
            assert e.acc == z["rec_acc"][r, i] == 0.0
            assert np.array_equal(np.array(e.obst, dtype=np.float64), z["rec_obst"][r, i])
        for k in range(meta["steps"]):
            rew, done, _ = vec.step([tuple(a) for a in z["rec_actions"][r, k]])
            assert [float(x) for x in rew] == list(z["rec_reward"][r, k]), (r, k)
            assert [int(d) for d in done] == list(z["rec_done"][r, k]), (r, k)


def test_features20_against_the_reference_helpers():
    """featureExtractor.py's numpy helpers (run through the shim) vs the oracle's restatement, 160 states."""
    from oracle.ballenv_oracle import features20
    z, meta = load_golden("features_kat")
    assert z["features"].shape == (meta["n"], 20)
    for i in range(meta["n"]):
        got = features20(tuple(z["agent"][i]), tuple(z["goal"][i]), [tuple(o) for o in z["obst"][i]],
                         agent_rad=float(z["agent_rad"][i]), obstacle_rad=meta["obstacle_rad"])
        np.testing.assert_allclose(got, z["features"][i], rtol=1e-12, atol=1e-12, err_msg=str(i))
    assert z["features"][:, 17:].max() > 1 and set(np.unique(z["features"][:, 0])) >= {0.0, 5.0}


def test_blocks29_against_the_reference_prep_state2():
    """The legacy 29-float block-count observation: prep_state2 of examples/ball_env_reinforce.py (lifted by the shim,
    tests/golden/blocks_kat.npz) vs the oracle's restatement, 200 states incl. dx == 0, dy == 0 and the block edges."""
    from oracle.ballenv_oracle import blocks29
    z, meta = load_golden("blocks_kat")
    assert z["blocks"].shape == (meta["n"], 29)
    for i in range(meta["n"]):
        got = blocks29(tuple(z["agent"][i]), tuple(z["goal"][i]), [tuple(o) for o in z["obst"][i]])
        assert np.array_equal(np.asarray(got), z["blocks"][i]), i
    assert np.all(z["blocks"][:, :4].sum(1) == 1) and z["blocks"][:, 16].min() >= 1 and z["blocks"][:, 4:].max() >= 3
