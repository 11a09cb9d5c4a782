"""ORACLE helper: structured synthetic (segmentation, embedding) streams for the clustering step.

They are built to reach every branch of ``OnlineSpeakerClustering.identify`` (reference
``src/diart/blocks/clustering.py:119-210``): first call, no active speaker, new centre, centre
update, ``cost >= delta`` un-mapping, forced assignment when no free centre is left
(small ``max_speakers``), short (non-``long``) speakers, NaN embedding rows and unused all-zero
centre rows (scipy's cosine ``cdist`` returns NaN for those).
"""
from __future__ import annotations

import numpy as np


def make_stream(n_chunks: int, seed: int, K: int = 3, D: int = 512, F: int = 293,
                n_protos: int = 8, sigma: float = 1.2, p_nan: float = 0.02,
                p_silent: float = 0.05):
    """-> seg (N,F,K) float32 in (0,1), emb (N,K,D) float32 (L2-normalised, NaN rows possible)."""
    rng = np.random.default_rng(seed)
    protos = rng.standard_normal((n_protos, D))
    protos /= np.linalg.norm(protos, axis=1, keepdims=True)
    seg = np.zeros((n_chunks, F, K), dtype=np.float32)
    emb = np.zeros((n_chunks, K, D), dtype=np.float32)
    t = np.linspace(0, 1, F)[:, None]
    for i in range(n_chunks):
        spk = rng.choice(n_protos, size=K, replace=False)
        kind = rng.choice(4, size=K, p=[0.25, 0.2, 0.45, 0.1])  # 0 inactive, 1 short, 2 long, 3 borderline
        if rng.random() < p_silent:
            kind[:] = 0
        base = 0.05 + 0.1 * rng.random((F, K))
        for k in range(K):
            if kind[k] == 1:      # a short burst above tau
                c, w = rng.uniform(0.1, 0.9), rng.uniform(0.02, 0.08)
                base[:, k] += 0.85 * np.exp(-0.5 * ((t[:, 0] - c) / w) ** 2)
            elif kind[k] == 2:    # long turn
                a, b = sorted(rng.uniform(0, 1, 2))
                b = max(b, a + 0.45)
                base[:, k] += 0.8 * ((t[:, 0] >= a) & (t[:, 0] <= b))
            elif kind[k] == 3:    # hovering around the thresholds
                base[:, k] += rng.uniform(0.2, 0.6)
        seg[i] = np.clip(base, 1e-4, 1 - 1e-4).astype(np.float32)
        e = protos[spk] + sigma / np.sqrt(D) * rng.standard_normal((K, D))
        e /= np.linalg.norm(e, axis=1, keepdims=True)
        emb[i] = e.astype(np.float32)
        for k in range(K):
            if rng.random() < p_nan:
                emb[i, k, :] = np.nan
    return seg, emb
