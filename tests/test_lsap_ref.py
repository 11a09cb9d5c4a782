"""The column-parallel assignment algorithm of csrc/cluster.cu (restated in oracle/lsap_ref.py)
against scipy.optimize.linear_sum_assignment, including heavy ties (1e10 blocks, zeros, integers)."""
import numpy as np
from scipy.optimize import linear_sum_assignment

from oracle.lsap_ref import lsap_rows


def test_lsap_matches_scipy_with_ties():
    rng = np.random.default_rng(0)
    for trial in range(4000):
        nr = int(rng.integers(1, 6))
        nc = int(rng.integers(nr, 24))
        mode = trial % 4
        if mode == 0:
            c = rng.random((nr, nc)) * 2
        elif mode == 1:
            c = rng.integers(0, 3, (nr, nc)).astype(float)
        elif mode == 2:
            c = rng.random((nr, nc)) * 2
            c[rng.random(nr) < 0.3, :] = 1e10
            c[:, rng.random(nc) < 0.5] = 1e10
            for _ in range(int(rng.integers(0, 3))):
                c[rng.integers(nr), rng.integers(nc)] = 0.0
        else:
            c = np.round(rng.random((nr, nc)) * 4) / 4
            c[:, rng.random(nc) < 0.4] = 1e10
        assert lsap_rows(c) == list(linear_sum_assignment(c)[1]), (trial, c)
