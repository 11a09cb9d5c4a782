"""ORACLE (test infrastructure): numpy restatement of the shared-identity merge rule implemented by
``cluster_export_kernel`` / ``cluster_merge_kernel`` (diart_b200/csrc/cluster.cu).  This mode is an extension beyond
the reference (BASELINE config 5): the reference has no multi-stream identity sharing, so the
oracle here is the single-process simulation of the same deterministic rule, applied in rank order.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np


def export_record(centers: np.ndarray, active: set, base: np.ndarray, base_active: set):
    """-> (payload (M,D) float64, kind (M,) in {0,1,2})"""
    M, D = centers.shape
    payload, kind = np.zeros((M, D)), np.zeros(M)
    for g in range(M):
        if g not in active:
            continue
        if g in base_active:
            payload[g], kind[g] = centers[g] - base[g], 1
        else:
            payload[g], kind[g] = centers[g], 2
    return payload, kind


def _cosine(c: np.ndarray, a: np.ndarray) -> float:
    cosv = float(np.dot(c, a)) / (np.sqrt(float(np.dot(c, c))) * np.sqrt(float(np.dot(a, a))))
    if abs(cosv) > 1.0:
        cosv = np.copysign(1.0, cosv)
    return 1.0 - cosv


def merge(records: List[Tuple[np.ndarray, np.ndarray]], base: np.ndarray, base_active: set, delta_new: float):
    """records[r] = (payload, kind) of rank r.  -> (centers, active, relabel[r] (M,) int) identical on every rank"""
    M, _ = base.shape
    centers, active = base.copy(), set(base_active)
    for payload, kind in records:                       # 1. deltas of existing centres, rank order
        for g in range(M):
            if kind[g] == 1:
                centers[g] += payload[g]
    relabels = [np.arange(M, dtype=np.int32) for _ in records]
    fresh = set()                                        # centres accepted during THIS merge
    for r, (payload, kind) in enumerate(records):       # 2. created centres, (rank, index) order
        used = set()                                     # targets of this rank's own creations stay distinct
        for g in range(M):
            if kind[g] != 2:
                continue
            c = payload[g]
            free = [a for a in range(M) if a not in active]
            # a rank already judged its creation to be new w.r.t. every centre it could see; only centres created
            # by EARLIER ranks in this step are candidates for being the same speaker (all centres if the table is full)
            cand = [a for a in sorted(fresh if free else active) if a not in used]
            best, bd = -1, np.inf
            for a in cand:
                d = _cosine(c, centers[a])
                if d < bd:
                    best, bd = a, d
            if best >= 0 and (bd < delta_new or not free):
                target = best
                centers[target] += c
            elif free:
                target = free[0]
                centers[target] = c
                fresh.add(target)
            else:                                        # full table and every centre already used by this rank
                continue
            active.add(target)
            used.add(target)
            relabels[r][g] = target
    return centers, active, relabels
