"""Host logic of the exact sampler (qsv/sampling.py): the level plan of the radix descent, emulated with Python
integers over virtual ranks, against a brute-force bisect over the exact cumulative weights in output order
(the random.choices rule of main/circuit.py:360-364) and against the oracle's float rule."""
import random

import numpy as np
import pytest

import helpers
import statevec_oracle as so
from qsv import sampling


def _random_weights(rng, n, kind):
    N = 1 << n
    if kind == "dense":
        w = rng.random(N)
    elif kind == "sparse":
        w = np.zeros(N)
        idx = rng.choice(N, size=max(1, N // 37), replace=False)
        w[idx] = rng.random(len(idx))
    else:                                   # a few heavy outcomes among tiny ones: many levels see ties in the high bits
        w = rng.random(N) * 1e-9
        w[rng.choice(N, size=3, replace=False)] = rng.random(3)
    return w / w.sum()


@pytest.mark.parametrize("n,g", [(5, 0), (13, 0), (14, 1), (15, 2), (16, 3), (17, 0)])
def test_descent_matches_bruteforce_in_output_order(n, g):
    rng = np.random.default_rng(100 * n + g)
    pyr = random.Random(n * 7 + g)
    for trial in range(4):
        src = list(range(n))
        if trial:
            pyr.shuffle(src)                # arbitrary output wiring: global positions anywhere in the output index
        w = _random_weights(rng, n, ("dense", "sparse", "spiky", "dense")[trial])
        for u in (0.0, pyr.random(), pyr.random(), 1.0 - 2.0 ** -53):
            want = helpers.exact_draw_bruteforce(w, n, src, u)
            got = helpers.exact_draw_emulated(w, n, g, src, u)
            assert got == want, (n, g, trial, u, got, want)


@pytest.mark.parametrize("n,g,k", [(13, 0, 3), (15, 2, 5), (16, 1, 12), (17, 0, 7)])
def test_descent_with_a_presummed_first_level(n, g, k):
    """The first level forced to the top k output bits (their bucket sums come from the last gate pass,
    qsv_run_pass_rounds_sums): the descent below them finds the same outcome as the plain plan and as brute force."""
    rng = np.random.default_rng(7 * n + k)
    pyr = random.Random(n + k)
    for trial in range(3):
        src = list(range(n))
        if trial:
            pyr.shuffle(src)
        w = _random_weights(rng, n, ("dense", "sparse", "spiky")[trial])
        steps = sampling.level_plan(src, n, n - g, first_bits=k)
        assert steps[0][1] == list(range(n - k, n)) and sum(len(s_[1]) for s_ in steps) == n
        for u in (0.0, pyr.random(), 1.0 - 2.0 ** -53):
            want = helpers.exact_draw_bruteforce(w, n, src, u)
            assert helpers.exact_draw_emulated(w, n, g, src, u, first_bits=k) == want
    tile = [0, 1, 2, 3, 4, 5, 9, 11, 12]
    assert sampling.presum_bits(None, 16, tile) == 10                      # 12, 11, 9 inside; stops at bit 5 (the fourth)
    assert sampling.presum_bits(list(range(15, -1, -1)), 16, tile) == 3     # the top output bits sit on positions 0, 1, 2
    assert sampling.presum_bits(None, 30, list(range(12))) == 12           # never more than one level's worth


def test_plan_shapes():
    # identity wiring at 30 qubits: 12 + (8 high + 2 low) bits, then levels of two low bits until <= 12 bits are left
    steps = sampling.level_plan(None, 30, 30)
    assert [s[0] for s in steps][-1] == "leaf"
    assert sum(len(s[1]) for s in steps) == 30
    assert steps[0][1] == list(range(18, 30)) and steps[0][3] == 0
    for kind, out_bits, pos, fixed in steps:
        assert len(out_bits) <= 12 and not any((fixed >> p) & 1 for p in pos)
    # worst case: the output index starts with the lowest physical positions
    src = list(range(29, -1, -1))
    steps = sampling.level_plan(src, 30, 27)
    assert sum(len(s[1]) for s in steps) == 30 and len(steps) <= 12
    with pytest.raises(RuntimeError):
        sampling.level_plan([0, 0, 1], 3, 3)


def test_u_fixed_is_exact():
    from fractions import Fraction
    pyr = random.Random(5)
    for u in [0.0, 0.5, 2.0 ** -60, 1.0 - 2.0 ** -53] + [pyr.random() for _ in range(50)]:
        m, s = sampling.u_fixed(u)
        assert m < (1 << 53) and Fraction(m, 1 << s) == Fraction(u)
    with pytest.raises(ValueError):
        sampling.u_fixed(1.0)


def test_exact_rule_agrees_with_the_float_rule_away_from_ties():
    """The reference's float accumulate + bisect (oracle.sample_index) and the exact rule pick the same outcome unless
    u * total lies within a rounding error of a cumulative sum - never the case for these draws."""
    rng = np.random.default_rng(3)
    pyr = random.Random(11)
    n = 10
    for _ in range(20):
        w = _random_weights(rng, n, "dense")
        u = pyr.random()
        assert helpers.exact_draw_bruteforce(w, n, None, u) == so.sample_index(list(w), u)


def test_zero_total_raises():
    with pytest.raises(ValueError):
        helpers.exact_draw_emulated(np.zeros(1 << 6), 6, 1, None, 0.3)
