"""ORACLE (test infrastructure, never shipped, never timed as the product).

numpy/float64 restatement of diart's constrained incremental clustering, i.e. of
``OnlineSpeakerClustering.identify/__call__`` (reference ``src/diart/blocks/clustering.py:119-218``)
together with the parts of ``SpeakerMap`` it reaches (``src/diart/mapping.py:179-360``):
every map mutation builds a NEW cost matrix and the Hungarian assignment is re-solved lazily on
the new matrix (``mapping.py:193-199,245-294``).  Here a "map" is just the (K, M) float64 cost
matrix; ``_solve`` is the lazily-evaluated ``_raw_optimal_assignments``.

Pinned against the reference's own code (imported by ``oracle/ref_import.py``) by
``tests/test_oracle_clustering.py`` and against the committed traces in ``tests/golden/``.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np
from scipy.optimize import linear_sum_assignment
from scipy.spatial.distance import cdist

INVALID = 1e10  # mapping.py:48-52 (MinimizationObjective.invalid_value)


def _solve(cost: np.ndarray) -> List[int]:
    """mapping.py:14-15 -- column chosen for every row (rows <= cols here)."""
    return list(linear_sum_assignment(cost, False)[1])


def _mapped_rows(cost: np.ndarray) -> List[int]:
    """mapping.py:17-21,185-187 -- rows whose best value is not the invalid marker."""
    return list(np.where(np.min(cost, axis=1) != INVALID)[0])


def _valid_assignments(cost: np.ndarray) -> Tuple[List[int], List[int]]:
    """mapping.py:214-231 with strict=False: keep (row, lsap col) iff the row is mapped."""
    mapped = _mapped_rows(cost)
    src, tgt = [], []
    for s, t in enumerate(_solve(cost)):
        if s in mapped:
            src.append(s)
            tgt.append(int(t))
    return src, tgt


class OracleClustering:
    """State + step function; attribute names follow the reference class (clustering.py:31-46)."""

    def __init__(self, tau_active: float, rho_update: float, delta_new: float,
                 metric: str = "cosine", max_speakers: int = 20):
        self.tau_active, self.rho_update, self.delta_new = tau_active, rho_update, delta_new
        self.metric, self.max_speakers = metric, max_speakers
        self.centers: Optional[np.ndarray] = None
        self.active_centers: set = set()
        self.blocked_centers: set = set()      # clustering.py:46 -- never populated in the reference
        self.last_margin = np.inf              # smallest |quantity - threshold| seen in the last step

    # clustering.py:48-71
    def _num_free(self) -> int:
        return self.max_speakers - len(self.active_centers) - len(self.blocked_centers)

    def _inactive(self) -> List[int]:
        return [c for c in range(self.max_speakers)
                if c not in self.active_centers or c in self.blocked_centers]

    def _next_free(self) -> Optional[int]:
        for c in range(self.max_speakers):
            if c not in self.active_centers and c not in self.blocked_centers:
                return c
        return None

    def _add_center(self, e: np.ndarray) -> int:
        c = self._next_free()                  # clustering.py:115-118
        self.centers[c] = e
        self.active_centers.add(c)
        return c

    def identify(self, seg: np.ndarray, emb: np.ndarray) -> np.ndarray:
        """seg (F,K) float32, emb (K,D) float32 -> final (K,M) float64 cost matrix."""
        emb = np.asarray(emb)
        K, M = seg.shape[1], self.max_speakers
        seg_max, seg_mean = np.max(seg, axis=0), np.mean(seg, axis=0)
        margin = min(np.min(np.abs(seg_max.astype(np.float64) - self.tau_active)),
                     np.min(np.abs(seg_mean.astype(np.float64) - self.rho_update)))
        active = np.where(seg_max >= self.tau_active)[0]            # :137-139
        long_spk = np.where(seg_mean >= self.rho_update)[0]         # :140-142
        no_nan = np.where(~np.isnan(emb).any(axis=1))[0]            # :144
        active = np.intersect1d(active, no_nan)                     # :145

        if self.centers is None:                                    # :149-158
            self.centers = np.zeros((M, emb.shape[1]))
            self.active_centers, self.blocked_centers = set(), set()
            cost = np.ones((K, M)) * INVALID
            for k in active:
                cost[k, self._add_center(emb[k])] = 0.0
            self.last_margin = margin
            return cost

        dist = cdist(emb, self.centers, metric=self.metric)         # :161, mapping.py:175
        inactive_rows = [k for k in range(K) if k not in active]    # :163-165
        cost = dist.copy()                                          # :166 / mapping.py:275-294
        for r in inactive_rows:
            cost[r, :] = INVALID
        for c in self._inactive():
            cost[:, c] = INVALID
        dist_map = cost
        # :168 / mapping.py:260-273 -- rows whose ASSIGNED cost is >= delta are unmapped entirely
        s, t = _valid_assignments(dist_map)
        valid = dist_map.copy()
        for r, c in zip(s, t):
            margin = min(margin, abs(dist_map[r, c] - self.delta_new)) if dist_map[r, c] != INVALID else margin
            if dist_map[r, c] >= self.delta_new:
                valid[r, :] = INVALID
        missed = [k for k in active if k not in _mapped_rows(valid)]  # :171-173

        new_center_speakers: List[int] = []
        for k in missed:                                            # :176-194
            has_space = len(new_center_speakers) < self._num_free()
            if has_space and k in long_spk:
                new_center_speakers.append(k)
            else:
                prefs = [g for g in np.argsort(dist_map[k, :]) if g in self.active_centers]
                _, taken = _valid_assignments(valid)
                free = [g for g in prefs if g not in taken]
                if free:
                    valid = valid.copy()
                    valid[k, free[0]] = 0.0                         # mapping.py:245-251
        s, t = _valid_assignments(valid)                            # :197-202
        for k, g in zip(s, t):
            if k not in missed and k in long_spk:
                assert g in self.active_centers, "Cannot update unknown centers"  # :98
                self.centers[g] += emb[k]
        for k in new_center_speakers:                               # :205-208
            valid = valid.copy()
            valid[k, self._add_center(emb[k])] = 0.0
        self.last_margin = margin
        return valid

    def __call__(self, seg: np.ndarray, emb: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        """-> (local->global map int32 (K,) with -1 for unmapped, permuted scores (F,M) float64).
        mapping.py:341-360 (``apply``)."""
        cost = self.identify(seg, emb)
        out = np.zeros((seg.shape[0], self.max_speakers))
        amap = -np.ones(seg.shape[1], dtype=np.int32)
        for k, g in zip(*_valid_assignments(cost)):
            out[:, g] = seg[:, k]
            amap[k] = g
        return amap, out
