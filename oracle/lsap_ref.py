"""ORACLE helper: Python restatement of the rectangular assignment algorithm (Crouse 2016, shortest
augmenting paths) in the column-parallel form that csrc/cluster.cu executes on one warp, including
scipy's tie-breaking (prefer an unassigned column among equal shortest-path costs; scan order = the
swap-removed `remaining` list that starts reversed).  Fuzzed against
``scipy.optimize.linear_sum_assignment`` by tests/test_lsap_ref.py."""
from __future__ import annotations

import numpy as np

INF = float("inf")


def lsap_rows(cost: np.ndarray):
    """cost (nr <= nc) -> column chosen for every row"""
    nr, nc = cost.shape
    assert nr <= nc
    u, v = [0.0] * nr, [0.0] * nc
    col4row, row4col = [-1] * nr, [-1] * nc
    for cur in range(nr):
        min_val, i = 0.0, cur
        pos = [nc - 1 - j for j in range(nc)]
        rem, spc, path = [True] * nc, [INF] * nc, [-1] * nc
        in_sc, in_sr = [False] * nc, [False] * nr
        num, sink = nc, -1
        while sink == -1:
            in_sr[i] = True
            for j in range(nc):                      # every lane at once on the GPU
                if rem[j]:
                    r = ((min_val + cost[i, j]) - u[i]) - v[j]
                    if r < spc[j]:
                        path[j], spc[j] = i, r
            lowest = min(spc[j] if rem[j] else INF for j in range(nc))
            cand = [j for j in range(nc) if rem[j] and spc[j] == lowest]
            free = [j for j in cand if row4col[j] == -1]
            j = max(free, key=lambda c: pos[c]) if free else min(cand, key=lambda c: pos[c])
            min_val = lowest
            if row4col[j] == -1:
                sink = j
            else:
                i = row4col[j]
            hole = pos[j]
            in_sc[j], rem[j] = True, False
            for c in range(nc):
                if c != j and rem[c] and pos[c] == num - 1:
                    pos[c] = hole
            num -= 1
        u[cur] += min_val
        for q in range(nr):
            if in_sr[q] and q != cur:
                u[q] += min_val - spc[col4row[q]]
        for j in range(nc):
            if in_sc[j]:
                v[j] -= min_val - spc[j]
        j = sink
        while True:
            q = path[j]
            row4col[j] = q
            col4row[q], j = j, col4row[q]
            if q == cur:
                break
    return col4row
