"""Random-draw addressing shared by the CPU oracle and the reference shim.

TEST INFRASTRUCTURE ONLY.  Nothing under ``gym_ballenv_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU
baseline legs use it (as the checker, never as the thing measured or shipped).

Why this exists
---------------
The reference draws everything from the process-global ``np.random`` MT19937
stream, in textual order (``gym_ballenv/envs/ballenv_env.py:24-25,45-46,
115-118,123-124,332,340,345,352``).  A sequential stream cannot be reproduced
by N environments stepping in parallel, so the B200 path gives every draw an
*address* and derives the 32-bit word for that address either from
Philox4x32-10 (production) or from an injected tape (parity mode):

    step draws   : (STREAM_STEP,  g, tick,    j)           one word per moving
                   obstacle j; a second word is derived from / supplied with it
    reset draws  : (STREAM_RESET, g, episode, kind|item|attempt)

``g`` is the *global* environment id, so trajectories do not depend on how the
environments are sharded over GPUs.  ``randint(n)`` is ``mulhi(word, n)``.

The same addressing is implemented in ``gym_ballenv_b200/csrc/ballenv_rng.cuh``;
this file is its CPU restatement and the oracle for it.
"""
from __future__ import annotations

import numpy as np

M32 = 0xFFFFFFFF
PHILOX_M0 = 0xD2511F53
PHILOX_M1 = 0xCD9E8D57
PHILOX_W0 = 0x9E3779B9
PHILOX_W1 = 0xBB67AE85

STREAM_STEP = 1
STREAM_RESET = 2
STREAM_ACTION = 3   # synthetic action stream used by bench/tests

# reset "kind" field (top 4 bits of counter word 2)
RK_HEAD = 0      # block 0: goal_x, goal_y, agent_x, agent_y   (gym ruleset)
                 # pygame : block 0 = goal (2 ranf), block 1 = first agent draw
RK_STATIC = 1    # | i << 16 | attempt >> 1 ; words (attempt & 1) * 2 + {0, 1}
RK_DYNAMIC = 2   # | j >> 1                 ; words (j & 1) * 2 + {0, 1}
RK_AGENT_REDRAW = 3   # | attempt            ; words 0..3 (two ranf)
RK_FIXED_AGENT = 4    # resetFixedstate (pygame): | outer << 12 | inner ; words 0..3 (two ranf); inner 0 = first draw


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Philox4x32-10 (Salmon et al., SC'11 "Parallel random numbers: as easy as
    1, 2, 3").  Scalar Python ints in, 4-tuple of uint32 out."""
    for r in range(10):
        p0 = PHILOX_M0 * c0
        p1 = PHILOX_M1 * c2
        hi0, lo0 = p0 >> 32, p0 & M32
        hi1, lo1 = p1 >> 32, p1 & M32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & M32, lo1, (hi0 ^ c3 ^ k1) & M32, lo0
        k0 = (k0 + PHILOX_W0) & M32
        k1 = (k1 + PHILOX_W1) & M32
    return c0, c1, c2, c3


def philox4x32_10_np(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10 over numpy uint64-held 32-bit lanes."""
    c0 = np.asarray(c0, dtype=np.uint64)
    c1 = np.asarray(c1, dtype=np.uint64)
    c2 = np.asarray(c2, dtype=np.uint64)
    c3 = np.asarray(c3, dtype=np.uint64)
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = np.uint64(k0)
    k1 = np.uint64(k1)
    m = np.uint64(M32)
    s32 = np.uint64(32)
    for r in range(10):
        p0 = np.uint64(PHILOX_M0) * c0
        p1 = np.uint64(PHILOX_M1) * c2
        hi0, lo0 = p0 >> s32, p0 & m
        hi1, lo1 = p1 >> s32, p1 & m
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & m, lo1, (hi0 ^ c3 ^ k1) & m, lo0
        k0 = (k0 + np.uint64(PHILOX_W0)) & m
        k1 = (k1 + np.uint64(PHILOX_W1)) & m
    return c0, c1, c2, c3


def mulhi(word, n):
    """randint(n) from a 32-bit word: floor(word * n / 2**32)."""
    return (int(word) * int(n)) >> 32


def mullo(word, n):
    return (int(word) * int(n)) & M32


def word_for(value, n):
    """Smallest word w with mulhi(w, n) == value (used to turn recorded
    reference draws into tape words)."""
    w = -((-int(value) << 32) // int(n))          # ceil(value * 2**32 / n)
    assert 0 <= w <= M32 and mulhi(w, n) == value, (value, n, w)
    return w


def ranf_from_words(a, b):
    """NumPy legacy ``random_sample``: 53-bit double from two 32-bit words
    (numpy/random/_legacy: ``(a >> 5) * 67108864 + (b >> 6)) / 2**53``)."""
    return ((int(a) >> 5) * 67108864.0 + (int(b) >> 6)) / 9007199254740992.0


def words_for_ranf(x):
    """Inverse of :func:`ranf_from_words` for a recorded double in [0, 1)."""
    v = int(x * 9007199254740992.0)
    assert v / 9007199254740992.0 == x
    a, b = (v >> 26) << 5, (v & ((1 << 26) - 1)) << 6
    assert ranf_from_words(a, b) == x
    return a, b


def reset_block(kind, item=0, attempt=0):
    """Counter word 2 of a reset draw and the first word index inside the
    Philox block -> (c2, word0)."""
    if kind == RK_HEAD:
        return (RK_HEAD << 28) | item, 0
    if kind == RK_STATIC:
        return (RK_STATIC << 28) | (item << 16) | (attempt >> 1), (attempt & 1) * 2
    if kind == RK_DYNAMIC:
        return (RK_DYNAMIC << 28) | (item >> 1), (item & 1) * 2
    if kind == RK_AGENT_REDRAW:
        return (RK_AGENT_REDRAW << 28) | attempt, 0
    if kind == RK_FIXED_AGENT:
        return (RK_FIXED_AGENT << 28) | (item << 12) | attempt, 0
    raise ValueError(kind)


class PhiloxDraws:
    """Production draw source: word(address) = Philox4x32-10(counter, key)."""

    def __init__(self, seed):
        self.k0 = seed & M32
        self.k1 = (seed >> 32) & M32

    # one word per moving obstacle; 4 obstacles share a Philox block
    def step_word(self, g, tick, j):
        blk = philox4x32_10(g & M32, tick & M32, j >> 2, STREAM_STEP, self.k0, self.k1)
        return blk[j & 3]

    def step_word2(self, g, tick, j, w1, n1):
        """Second word of a step draw: the low half of w1 * n1 (the part of the
        product the first draw did not use)."""
        return mullo(w1, n1)

    def reset_words(self, g, episode, kind, item=0, attempt=0, count=2):
        c2, w0 = reset_block(kind, item, attempt)
        blk = philox4x32_10(g & M32, episode & M32, c2, STREAM_RESET, self.k0, self.k1)
        return blk[w0:w0 + count]

    def action_word(self, g, t):
        return philox4x32_10(g & M32, t & M32, 0, STREAM_ACTION, self.k0, self.k1)[0]


class TapeDraws:
    """Parity-mode draw source: words come from injected tapes.

    step_tape  : uint32 [T, n_envs, n_dynamic, 2]   indexed by (tick - tick0, g - g0, j, {w1, w2})
    reset_tape : uint32 [E, n_envs, R] with R = 4 + 2 * A * n_static + 2 * n_dynamic
                 slot layout: head 0..3 | static (i * A + attempt) * 2 + {0,1} | dynamic 2 j + {0,1}
    """

    def __init__(self, step_tape=None, reset_tape=None, n_static=0, n_dynamic=0,
                 attempts=1, g0=0, tick0=0, episode0=0):
        self.step_tape = step_tape
        self.reset_tape = reset_tape
        self.ks, self.kd, self.A = n_static, n_dynamic, attempts
        self.g0, self.tick0, self.episode0 = g0, tick0, episode0

    def step_word(self, g, tick, j):
        return int(self.step_tape[tick - self.tick0, g - self.g0, j, 0])

    def step_word2(self, g, tick, j, w1, n1):
        return int(self.step_tape[tick - self.tick0, g - self.g0, j, 1])

    def reset_words(self, g, episode, kind, item=0, attempt=0, count=2):
        row = self.reset_tape[episode - self.episode0, g - self.g0]
        if kind == RK_HEAD:
            s = 0
        elif kind == RK_STATIC:
            if attempt >= self.A:
                raise IndexError("reset tape holds %d attempts per static obstacle" % self.A)
            s = 4 + (item * self.A + attempt) * 2
        elif kind == RK_DYNAMIC:
            s = 4 + 2 * self.A * self.ks + 2 * item
        else:
            raise ValueError("tape mode has no slot for kind %d" % kind)
        return tuple(int(w) for w in row[s:s + count])


def reset_tape_width(n_static, n_dynamic, attempts):
    return 4 + 2 * attempts * n_static + 2 * n_dynamic
