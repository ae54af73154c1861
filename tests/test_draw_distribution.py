"""Distribution of the draw scheme: the kernels replace the reference's
``np.random.randint(n)`` calls (gym_ballenv/envs/ballenv_env.py:24-25, 115-118, 332, 340, 345, 352) by
``randint(n) = mulhi(word, n)`` on Philox4x32-10 words, and take an obstacle's SECOND draw of a step from the unused
low half of the first product (``w2 = w1 * 100 mod 2^32``, then ``mulhi(w2, 9)``).  Equality with the port is tested
elsewhere; here: are the values uniform over their ranges, and is the second draw independent of the first where the
reference uses it (only when the first draw is >= rd_th_obs, ballenv_env.py:332-342)?  Chi-square tests over 10^6
words with fixed seeds; the thresholds are the 99.9 % quantiles, so a correct scheme fails one run in a thousand
per test - with these seeds it passes."""
import numpy as np
import pytest
from scipy import stats

from oracle import draws as D

N = 1_000_000


def _words(seed, n=N, stream=D.STREAM_STEP):
    """n Philox words as the step kernel addresses them: counter (g, tick, quad, stream), word j & 3."""
    g = np.arange(n // 4, dtype=np.uint64)
    blk = D.philox4x32_10_np(g & np.uint64(0xffff), g >> np.uint64(16), np.uint64(0), np.uint64(stream), seed & D.M32,
                             (seed >> 32) & D.M32)
    return np.stack(blk, 1).reshape(-1).astype(np.uint64)


def _mulhi(w, n):
    return (w * np.uint64(n)) >> np.uint64(32)


def _chi2_uniform(values, k):
    counts = np.bincount(values.astype(np.int64), minlength=k)
    assert counts.size == k
    return stats.chisquare(counts).statistic, stats.chi2.ppf(0.999, k - 1)


@pytest.mark.parametrize("n", [100, 9, 4, 23, 500, 460, 20, 10])
def test_randint_is_uniform(n):
    """mulhi(word, n) over the ranges the environment draws from: 100 (goal-seeking test), 9 (move list), n_goals - 1
    (4 for the defaults, 23 for the dense configuration) and the reset ranges 500 / 460 / 20 / 10."""
    stat, crit = _chi2_uniform(_mulhi(_words(seed=12345 + n), n), n)
    assert stat < crit, (n, stat, crit)


def test_second_draw_is_uniform_and_independent_of_the_first():
    """(u, i) = (mulhi(w, 100), mulhi(w * 100 mod 2^32, 9)): i is uniform on [0, 9), and uniform within every value of
    u the reference can pair it with (u >= rd_th_obs = 60: the random-move branch) - a 40 x 9 contingency table -
    and also over the coarse split the kernel branches on (u < 60 | u >= 60)."""
    w = _words(seed=987654321)
    u = _mulhi(w, 100)
    w2 = (w * np.uint64(100)) & np.uint64(D.M32)
    i = _mulhi(w2, 9)
    stat, crit = _chi2_uniform(i, 9)
    assert stat < crit
    sel = u >= 60
    table = np.zeros((40, 9), dtype=np.int64)
    np.add.at(table, ((u[sel] - np.uint64(60)).astype(np.int64), i[sel].astype(np.int64)), 1)
    chi2, pval, dof, _ = stats.chi2_contingency(table)
    assert dof == 39 * 8 and chi2 < stats.chi2.ppf(0.999, dof), (chi2, dof)
    split = np.zeros((2, 9), dtype=np.int64)
    np.add.at(split, (sel.astype(np.int64), i.astype(np.int64)), 1)
    chi2, pval, dof, _ = stats.chi2_contingency(split)
    assert chi2 < stats.chi2.ppf(0.999, dof), (chi2, dof)
    # the joint cell counts against the product of the exact marginals (P(u >= 60) = 0.4, P(i) = 1 / 9)
    expected = np.array([[0.6 / 9] * 9, [0.4 / 9] * 9]) * w.size
    stat = ((split - expected) ** 2 / expected).sum()
    assert stat < stats.chi2.ppf(0.999, 17), stat


def test_scalar_and_vector_philox_agree_and_match_mulhi():
    """The vectorised generator used above is the scalar one the oracle steps with (and the kernels restate)."""
    blk = D.philox4x32_10_np(np.arange(8), 7, 2, D.STREAM_STEP, 0xdeadbeef, 0x1234)
    for g in range(8):
        assert tuple(int(b[g]) for b in blk) == D.philox4x32_10(g, 7, 2, D.STREAM_STEP, 0xdeadbeef, 0x1234)
    assert D.mulhi(0xffffffff, 100) == 99 and D.mulhi(0, 100) == 0 and D.mulhi(1 << 31, 9) == 4
