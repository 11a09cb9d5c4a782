"""Overlapping-window aggregation after the hot path (mirrors the behaviour of reference
``src/diart/blocks/aggregation.py:73-218``; SURVEY.md 8(f) "next" row 1).  Host-side numpy."""
from __future__ import annotations

from typing import List, Optional

import numpy as np

from ..core import Segment, SlidingWindow, SlidingWindowFeature

_MODES = ("strict", "loose", "center")


def _crop(buffer: SlidingWindowFeature, focus: Segment, mode: str) -> np.ndarray:
    return buffer.crop(focus, mode=mode, fixed=focus.duration)


def _hamming(buffers, focus, mode):
    """average weighted by the Hamming window aligned to each buffer (aggregation.py:73-92)"""
    num_frames = buffers[0].data.shape[0]
    window = np.expand_dims(np.hamming(num_frames), axis=-1)
    ham = np.stack([_crop(SlidingWindowFeature(window, b.sliding_window), focus, mode) for b in buffers])
    val = np.stack([_crop(b, focus, mode) for b in buffers])
    return np.sum(ham * val, axis=0) / np.sum(ham, axis=0)


def _mean(buffers, focus, mode):
    return np.mean(np.stack([_crop(b, focus, mode) for b in buffers]), axis=0)


def _first(buffers, focus, mode):
    return _crop(buffers[0], focus, mode)


_STRATEGIES = {"hamming": _hamming, "mean": _mean, "first": _first}


class DelayedAggregation:
    """Aggregates the ``latency / step`` most recent buffers over the region that ends ``latency``
    seconds before the newest buffer's end (aggregation.py:120-218)."""

    def __init__(self, step: float, latency: Optional[float] = None, strategy: str = "hamming",
                 cropping_mode: str = "loose"):
        assert cropping_mode in _MODES, f"Invalid cropping mode `{cropping_mode}`"
        assert strategy in _STRATEGIES
        self.step = step
        self.latency = step if latency is None else latency
        assert self.step <= self.latency, "Invalid latency requested"
        self.strategy, self.cropping_mode = strategy, cropping_mode
        self.num_overlapping_windows = int(round(self.latency / self.step))

    def __call__(self, buffers: List[SlidingWindowFeature]) -> SlidingWindowFeature:
        start = buffers[-1].extent.end - self.latency
        region = Segment(start, start + self.step)
        values = _STRATEGIES[self.strategy](buffers, region, self.cropping_mode)
        res = region.duration / values.shape[0]
        window = SlidingWindowFeature(values, SlidingWindow(start=region.start, duration=res, step=res))
        # very first buffer of a stream: also emit everything before the region (aggregation.py:188-212)
        if len(buffers) == 1 and buffers[-1].extent.start == 0:
            first_region = Segment(0, region.end)
            first = buffers[0].crop(first_region, mode=self.cropping_mode, fixed=first_region.duration)
            first[-values.shape[0]:] = values
            res = region.end / first.shape[0]
            window = SlidingWindowFeature(first, SlidingWindow(start=0, duration=res, step=res))
        return window
