"""Boundary plumbing for temporal features (mirrors reference ``src/diart/features.py:77-138``).

Blocks accept ``SlidingWindowFeature`` / ``numpy.ndarray`` / ``torch.Tensor`` of shape (frames, dim)
or (batch, frames, dim); ``cast`` turns any of them into a float32 tensor (batch, frames, dim) and
remembers what came in, ``restore_type`` gives results back in that form.
"""
from __future__ import annotations

from typing import Optional, Union

import numpy as np
import torch

from .core import SlidingWindow, SlidingWindowFeature

TemporalFeatures = Union[SlidingWindowFeature, np.ndarray, torch.Tensor]


class TemporalFeatureFormatter:
    def __init__(self):
        self._kind: Optional[str] = None
        self._duration = 0.0
        self._start = 0.0

    def cast(self, features: TemporalFeatures) -> torch.Tensor:
        if isinstance(features, SlidingWindowFeature):
            sw = features.sliding_window
            assert sw.duration == sw.step, "Features sliding window duration and step must be equal"
            self._kind, self._start = "swf", sw.start
            self._duration = features.data.shape[0] * sw.duration
            data = torch.from_numpy(features.data)
        elif isinstance(features, np.ndarray):
            self._kind, data = "numpy", torch.from_numpy(features)
        elif isinstance(features, torch.Tensor):
            self._kind, data = "torch", features
        else:
            raise ValueError("Unknown format. Provide one of SlidingWindowFeature, numpy.ndarray, torch.Tensor")
        assert data.ndim in (2, 3), "Temporal features must be 2D or 3D"
        if data.ndim == 2:
            data = data.unsqueeze(0)
        return data.float()

    def restore_type(self, features: torch.Tensor) -> TemporalFeatures:
        if self._kind == "swf":
            batch_size, num_frames, _ = features.shape
            assert batch_size == 1, "Batched SlidingWindowFeature objects are not supported"
            res = self._duration / num_frames
            return SlidingWindowFeature(features.squeeze(dim=0).cpu().numpy(),
                                        SlidingWindow(start=self._start, duration=res, step=res))
        if self._kind == "numpy":
            return features.cpu().numpy()
        return features
