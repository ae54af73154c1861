"""The C restatement of the oracle (oracle/ballenv_oracle.c) against the Python restatement, bit for bit.
The Python restatement is the one pinned to the reference's golden fixtures (test_oracle_vs_golden.py)."""
import numpy as np
import pytest

from oracle import draws as D
from oracle.ballenv_oracle import OracleConfig, OracleVec
from oracle.c_oracle import COracleVec

DENSE_GOALS = [(x, y) for y in (100, 200, 300, 400) for x in (50, 130, 210, 290, 370, 450)]


@pytest.mark.parametrize("cfg", [
    OracleConfig(window=5, max_episode_steps=13),
    OracleConfig(window=10, n_static=8, n_dynamic=24, speeds=[1] * 24, goals=DENSE_GOALS, max_episode_steps=9),
    OracleConfig(window=7, n_static=3, n_dynamic=2, speeds=[2, 3], goals=[(5, 5), (5, 5), (400, 300)], change_step=4,
                 max_episode_steps=0, auto_reset=False),
])
def test_c_oracle_equals_python_oracle(cfg):
    n, T, seed, g0 = 24, 60, 99, 1000
    py = OracleVec(cfg, D.PhiloxDraws(seed), n, g0)
    c = COracleVec(cfg, seed, n, g0)
    py.reset()
    c.reset()
    assert np.array_equal(c.observe(), np.array(py.observe(), dtype=np.float32))
    rng = np.random.RandomState(4)
    for t in range(T):
        a = rng.randint(0, 9, n)
        r_py, d_py, f_py = py.step(a.tolist())
        r_c, d_c, f_c = c.step(a)
        assert np.array_equal(r_c, np.array(r_py, dtype=np.float64)), t       # same fp64 operations: bit-exact
        assert np.array_equal(d_c, np.array(d_py)), t
        assert np.array_equal(f_c, np.array(f_py, dtype=np.uint8)), t
        assert np.array_equal(c.observe(), np.array(py.observe(), dtype=np.float32)), t
    st = c.state()
    for i, e in enumerate(py.envs):
        assert tuple(st["agent"][i]) == tuple(e.agent) and tuple(st["goal"][i]) == tuple(e.goal)
        assert st["dist"][i] == e.dist and st["total"][i] == e.total_distance and st["acc"][i] == e.acc
        assert np.array_equal(st["obstacles"][i], np.array(e.obst, dtype=np.float64).reshape(-1, 2))
        assert list(st["dyn_goal"][i]) == e.goal_idx and list(st["dyn_counter"][i]) == e.counter
        assert (st["ep_len"][i], st["episode"][i], st["tick"][i]) == (e.ep_len, e.episode, e.tick)
    for k, v in py.stats.items():
        assert c.stats[k] == v, k


def test_c_oracle_threads_do_not_change_results(monkeypatch):
    cfg = OracleConfig(window=5, max_episode_steps=20)
    outs = []
    for threads in ("1", "4"):
        monkeypatch.setenv("ORC_THREADS", threads)
        c = COracleVec(cfg, 3, 600)
        c.reset()
        rng = np.random.RandomState(1)
        for t in range(25):
            r, d, f = c.step(rng.randint(0, 9, 600))
        outs.append((c.observe(), r, d, f, c.stats))
    assert all(np.array_equal(a, b) for a, b in zip(outs[0][:4], outs[1][:4])) and outs[0][4] == outs[1][4]
