"""CPU tests of the host-side mirror of the reference interface (config schema, sharding arithmetic, bench byte model)."""
import os
import sys
from argparse import Namespace

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _args(**over):
    # read_arguments() defaults of examples/ball_cnn_ac3.py:37-59
    d = dict(static_obstacles=13, dynamic_obstacles=5, obstacle_speed=[1, 1, 1, 1, 1],
             obs_goal_position=['12,122', '123,93', '87,150', '430,440', '230,11'], time_step_for_change=50,
             rd_th_obs=60, rd_th_agent=80, static_thresholds=[0, 0], dynamic_thresholds=[10, 10],
             static_penalty=[1, 1], dynamic_penalty=[4000, 8000])
    d.update(over)
    return Namespace(**d)


def test_env_config_from_reference_namespace():
    from gym_ballenv_b200 import EnvConfig
    from gym_ballenv_b200 import _lib as L
    cfg = EnvConfig.from_args(_args())
    assert cfg.goals() == [(12, 122), (123, 93), (87, 150), (430, 440), (230, 11)]      # ballenv_env.py:99-103
    c = cfg.to_c(window=10)
    assert (c.static_obstacles, c.dynamic_obstacles, c.n_goals, c.window) == (13, 5, 5, 10)
    assert c.static_penalty == 1.0 and c.dynamic_penalty == 8000.0                      # only index 1 is used (:138,158)
    assert c.ruleset == L.RULESET_GYM and c.max_episode_steps == 1000


@pytest.mark.parametrize("over,msg", [
    (dict(obstacle_speed=[1, 1]), "obstacle_speed"),
    (dict(obs_goal_position=['1,2']), "obstacle_goal_position"),
    (dict(static_penalty=[1]), "static_penalty"),
    (dict(dynamic_thresholds=[1, 2, 3]), "dynamic_thresholds"),
])
def test_assert_arguments_mirror(over, msg):
    """Same list-length checks as assert_arguments() (examples/ball_cnn_ac3.py:61-68), as ValueError."""
    from gym_ballenv_b200 import EnvConfig
    with pytest.raises(ValueError, match=msg):
        EnvConfig.from_args(_args(**over))


def test_cli_string_speeds_are_accepted():
    """--obstacle_speed has no type= in the reference (examples/ball_cnn_ac3.py:42): CLI values arrive as strings."""
    from gym_ballenv_b200 import EnvConfig
    c = EnvConfig.from_args(_args(obstacle_speed=['2', '1', '1', '3', '1'])).to_c(window=5)
    assert [c.obstacle_speed[i] for i in range(5)] == [2.0, 1.0, 1.0, 3.0, 1.0]


def test_dense_moving_is_config_3():
    from gym_ballenv_b200 import EnvConfig
    cfg = EnvConfig.dense_moving()
    assert (cfg.static_obstacles, cfg.dynamic_obstacles, len(cfg.goals())) == (8, 24, 24)
    assert cfg.goals()[0] == (50, 100) and cfg.goals()[-1] == (450, 400)


def test_shard_bounds_cover_and_order():
    from gym_ballenv_b200.distributed import shard_bounds
    for total in (0, 1, 7, 65536, 1 << 20, 1000003):
        for world in (1, 2, 3, 4, 8):
            pos = 0
            for r in range(world):
                off, cnt = shard_bounds(total, r, world)
                assert off == pos and cnt in (total // world, total // world + 1)
                pos += cnt
            assert pos == total
    assert shard_bounds(1 << 20, 7, 8) == (7 * 131072, 131072)       # BASELINE.json config 4
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def test_bench_byte_model_matches_survey():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.alg_bytes_per_env_step(bench.workload_spec("w5")) == 405       # 204 + 201
    assert bench.alg_bytes_per_env_step(bench.workload_spec("c3")) == 1121      # 392 + 729
    io = 8 + 4 * 104 + 4 + 1
    assert bench.moved_bytes_per_env_step(bench.workload_spec("c3"), 200) == pytest.approx(io + (1121 - io) / 200)
    a = bench.config_dict(bench.workload_spec("c3"), 65536, 8, 200)
    assert a["total_envs"] == 8 * 65536 and "workload" in a and "l2" in a


def test_path_log_reader_roundtrip_and_reference_head(tmp_path):
    """gym_ballenv_b200.pathlogs: the reference's demonstration-log layout (lists of 29-float prep_state2 vectors and
    [+-10, +-10] actions) written and read back; labels as examples/train_supervise.py derives them; and, where the
    reference tree is present, the shipped Python 2 pickles themselves against the recorded head of the log."""
    import json
    import numpy as np
    import torch
    from gym_ballenv_b200 import pathlogs as P
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "pathlog_kat.npz"))
    meta = json.loads(str(z["meta"]))
    assert np.array_equal(P.action_labels(z["actions"]), z["labels"])
    assert P.action_labels(np.array([[20, 0], [0, 0]]))[0] == 8          # not in move_list: the search falls through
    sf, af = str(tmp_path / "states"), str(tmp_path / "actions")
    P.save_path_log(torch.from_numpy(z["states"]), torch.from_numpy(z["labels"]), sf, af)
    x, y = P.load_path_log(sf, af)
    assert x.dtype == torch.float32 and tuple(x.shape) == (meta["n"], 29)
    assert np.array_equal(x.numpy(), z["states"].astype(np.float32)) and np.array_equal(y.numpy(), z["labels"])
    # every logged vector is a prep_state2 output: one goal-quadrant bit, the agent's own cell counted
    assert np.all(z["states"][:, :4].sum(1) == 1) and np.all(z["states"][:, 16] >= 1)
    ref = "/root/reference/examples"
    if os.path.exists(os.path.join(ref, "State_info_trail_no2")):
        x, y = P.load_path_log(os.path.join(ref, "State_info_trail_no2"), os.path.join(ref, "Trial_no_2"))
        assert len(x) == meta["total"] == len(y)
        assert np.array_equal(x[:meta["n"]].numpy(), z["states"].astype(np.float32))
        assert np.array_equal(y[:meta["n"]].numpy(), z["labels"])
        assert np.bincount(y.numpy(), minlength=9).tolist() == meta["label_histogram"]


def test_legacy_block_policy_supervised_fit_and_reference_checkpoint():
    """gym_ballenv_b200.legacy: the 29 -> 128 -> 128 -> 9 net of the older scripts fits the head of the shipped
    demonstration log (the supervised loop of examples/train_supervise.py), and, where the reference tree is present,
    loads its supervised checkpoint by parameter name and labels that log better than chance."""
    import json
    import numpy as np
    import torch
    from gym_ballenv_b200 import pathlogs as P
    from gym_ballenv_b200.legacy import BlockPolicy, train_supervised
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "pathlog_kat.npz"))
    x, y = torch.from_numpy(z["states"]).float(), torch.from_numpy(z["labels"])
    torch.manual_seed(0)
    model = BlockPolicy(logits=True)
    first = train_supervised(model, x, y, epochs=1, batch_size=32)
    last = train_supervised(model, x, y, epochs=150, batch_size=32)
    assert last < first
    probs = BlockPolicy()(x)
    assert probs.shape == (len(x), 9) and torch.allclose(probs.sum(-1), torch.ones(len(x)), atol=1e-6)
    ckpt = "/root/reference/examples/stored_models/supervised/episode_9999.pth"
    if os.path.exists(ckpt):
        ref = BlockPolicy(logits=True)
        ref.load_state_dict(torch.load(ckpt, map_location="cpu"))
        xs, ys = P.load_path_log("/root/reference/examples/State_info_trail_no2", "/root/reference/examples/Trial_no_2")
        acc = (ref(xs).argmax(-1) == ys).float().mean().item()
        assert acc > 0.3, acc          # nine classes; the net was trained on logs like this one


def test_registers_the_reference_id_when_gym_is_present(monkeypatch):
    """gym_ballenv/__init__.py:4-11 registers 'gymball-v0'; with a gym importable this package registers the same id
    (entry point BallEnv, TimeLimit 1000) unless the id is taken or BALLENV_NO_GYM_REGISTER=1."""
    import types
    import gym_ballenv_b200 as pkg
    calls = []

    class GymError(Exception):
        pass

    gym = types.ModuleType("gym")
    gym.error = types.SimpleNamespace(Error=GymError)
    envs = types.ModuleType("gym.envs")
    reg = types.ModuleType("gym.envs.registration")

    def register(**kw):
        if any(c["id"] == kw["id"] for c in calls):
            raise GymError("Cannot re-register id")
        calls.append(kw)

    reg.register = register
    gym.envs, envs.registration = envs, reg
    monkeypatch.setitem(sys.modules, "gym", gym)
    monkeypatch.setitem(sys.modules, "gym.envs", envs)
    monkeypatch.setitem(sys.modules, "gym.envs.registration", reg)
    assert pkg._register_with_gym() is True
    assert calls == [dict(id='gymball-v0', entry_point='gym_ballenv_b200.env:BallEnv', max_episode_steps=1000,
                          reward_threshold=100.0, nondeterministic=False)]
    assert pkg._register_with_gym() is False        # taken: left alone, no exception
    monkeypatch.setenv("BALLENV_NO_GYM_REGISTER", "1")
    calls.clear()
    assert pkg._register_with_gym() is False and calls == []
