"""``Binarize`` (mirrors reference ``src/diart/blocks/utils.py:11-59``): discrete-time scores ->
speaker turns at frame middles, label ``speaker<g>``."""
from __future__ import annotations

from typing import Optional

import numpy as np

from ..core import Annotation, Segment, SlidingWindowFeature


class Binarize:
    def __init__(self, threshold: float, uri: Optional[str] = None):
        self.uri = uri
        self.threshold = threshold

    def __call__(self, segmentation: SlidingWindowFeature) -> Annotation:
        num_frames, num_speakers = segmentation.data.shape
        grid = segmentation.sliding_window
        active = segmentation.data > self.threshold
        annotation = Annotation(uri=self.uri, modality="speech")
        middles = grid.start + grid.step * np.arange(num_frames + 1) + 0.5 * grid.duration
        for spk in np.where(active.any(axis=0))[0]:
            col = np.concatenate([[False], active[:, spk], [False]])
            change = np.flatnonzero(col[1:] != col[:-1])       # on/off boundaries, in frame units
            # a turn that is active from frame 0 starts at the first frame's middle
            for on, off in zip(change[0::2], change[1::2]):
                annotation[Segment(middles[on], middles[off]), int(spk)] = f"speaker{spk}"
        return annotation
