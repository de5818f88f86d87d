"""Pins oracle/lsap_oracle.c (C restatement of the published Crouse/JV solve)
against scipy.optimize.linear_sum_assignment, the reference's own solver
(sedt/matcher.py:95): bit-exact indices, including ties and degenerate shapes."""
import numpy as np
import pytest
from scipy.optimize import linear_sum_assignment

from oracle import matcher_oracle


@pytest.mark.parametrize("nr,nc", [(20, 10), (20, 0), (20, 1), (20, 20), (20, 28), (10, 12), (1, 1), (3, 7), (31, 5)])
def test_lsap_c_random(nr, nc):
    rng = np.random.default_rng(nr * 100 + nc)
    for _ in range(50):
        C = rng.standard_normal((nr, nc)).astype(np.float32)
        a, b = matcher_oracle.lsap_c(C)
        r, c = linear_sum_assignment(C)
        assert np.array_equal(a, r) and np.array_equal(b, c)


@pytest.mark.parametrize("nr,nc", [(20, 10), (8, 8), (6, 15), (20, 28)])
def test_lsap_c_ties(nr, nc):
    rng = np.random.default_rng(7)
    for hi in (1, 2, 4):
        for _ in range(100):
            C = rng.integers(0, hi + 1, size=(nr, nc)).astype(np.float32)   # heavy ties
            a, b = matcher_oracle.lsap_c(C)
            r, c = linear_sum_assignment(C)
            assert np.array_equal(a, r) and np.array_equal(b, c)
    C = np.ones((nr, nc), np.float32)                                      # constant matrix
    a, b = matcher_oracle.lsap_c(C)
    r, c = linear_sum_assignment(C)
    assert np.array_equal(a, r) and np.array_equal(b, c)


def test_lsap_c_invalid():
    C = np.zeros((3, 3)); C[1, 1] = np.nan
    with pytest.raises(ValueError):
        matcher_oracle.lsap_c(C)
    with pytest.raises(ValueError):
        linear_sum_assignment(C)
    C = np.full((3, 3), np.inf)
    with pytest.raises(ValueError):
        matcher_oracle.lsap_c(C)


def test_lsap_c_batched():
    rng = np.random.default_rng(1)
    B, Q, ldk = 40, 20, 28
    cost = rng.standard_normal((B, Q, ldk)).astype(np.float32)
    K = rng.integers(0, ldk + 1, size=B).astype(np.int32)
    rows, cols, counts = matcher_oracle.lsap_c_batched(cost, K)
    for b in range(B):
        r, c = linear_sum_assignment(cost[b, :, :K[b]])
        assert counts[b] == len(r)
        assert np.array_equal(rows[b, :counts[b]], r) and np.array_equal(cols[b, :counts[b]], c)
