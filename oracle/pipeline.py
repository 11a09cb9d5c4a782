"""ORACLE (test infrastructure / CPU baseline; never the product path).

CPU restatement of the reference's per-batch hot path, ``SpeakerDiarization.__call__`` lines 177-203
(reference ``src/diart/blocks/diarization.py``): segmentation -> overlapped-speech penalty ->
embedding (waveform repeated once per local speaker, ``blocks/embedding.py:57-59``) -> L2
normalisation -> the sequential clustering loop.  Networks: ``oracle/nets.py`` (torch, CPU);
clustering: ``oracle/clustering.py``.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np
import torch

from .clustering import OracleClustering


def overlapped_speech_penalty(seg: torch.Tensor, gamma: float = 3, beta: float = 10) -> torch.Tensor:
    """reference ``src/diart/functional.py:6-13``"""
    probs = torch.softmax(beta * seg, dim=-1)
    weights = torch.pow(seg, gamma) * torch.pow(probs, gamma)
    weights[weights < 1e-8] = 1e-8
    return weights


def osp_block(seg: torch.Tensor, gamma: float = 3, beta: float = 10, normalize: bool = False) -> torch.Tensor:
    """reference ``src/diart/blocks/embedding.py:98-107``"""
    w = overlapped_speech_penalty(seg, gamma, beta)
    if normalize:
        lo = w.min(dim=1, keepdim=True).values
        hi = w.max(dim=1, keepdim=True).values
        w = (w - lo) / (hi - lo)
        w.nan_to_num_(1e-8)
    return w


def normalize_embeddings(emb: torch.Tensor, norm: float = 1) -> torch.Tensor:
    """reference ``src/diart/functional.py:16-27``"""
    if emb.ndim == 2:
        emb = emb.unsqueeze(0)
    return norm * emb / torch.norm(emb, p=2, dim=-1, keepdim=True)


class OraclePipeline:
    def __init__(self, seg_net, emb_net, tau_active=0.6, rho_update=0.3, delta_new=1.0, gamma=3, beta=10,
                 max_speakers=20, normalize_weights=False, as_reference: bool = True):
        self.seg_net, self.emb_net = seg_net, emb_net
        self.gamma, self.beta, self.normalize_weights = gamma, beta, normalize_weights
        self.as_reference = as_reference     # True: K-fold repeated trunk exactly like the reference
        self.clustering = OracleClustering(tau_active, rho_update, delta_new, "cosine", max_speakers)

    @torch.no_grad()
    def nets(self, batch: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """batch (B,S) -> segmentation (B,F,K), unit-norm embeddings (B,K,D)"""
        wave = batch[:, None, :]
        seg = self.seg_net(wave)
        w = osp_block(seg, self.gamma, self.beta, self.normalize_weights)
        B, F, K = w.shape
        if self.as_reference:
            rep = wave.repeat(1, K, 1).reshape(B * K, 1, -1)
            emb = self.emb_net(rep, w.permute(0, 2, 1).reshape(B * K, F)).reshape(B, K, -1)
        else:
            emb = self.emb_net.forward_dedup(wave, w)
        return seg, normalize_embeddings(emb)

    def __call__(self, batch: torch.Tensor):
        """-> seg (B,F,K) f32, emb (B,K,D) f32, maps (B,K) int32, smallest decision margin per chunk"""
        seg, emb = self.nets(batch)
        seg_np, emb_np = seg.numpy(), emb.numpy()
        maps, margins = [], []
        for s, e in zip(seg_np, emb_np):
            amap, _ = self.clustering(s, e)
            maps.append(amap)
            margins.append(self.clustering.last_margin)
        return seg_np, emb_np, np.stack(maps), np.array(margins)
