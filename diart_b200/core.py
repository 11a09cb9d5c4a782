"""Minimal stand-in for the parts of ``pyannote.core`` that diart's hot path touches.

diart types its block interfaces with ``pyannote.core`` objects
(``SlidingWindowFeature`` in/out of ``SpeakerDiarization.__call__``,
reference ``src/diart/blocks/diarization.py:157-203``; ``features.py:8``).
``pyannote.core`` is not installed in this image, so when the real package is
importable we re-export it, otherwise we provide the small surface listed in
SURVEY.md Appendix B.  The same objects are used on the oracle side and on the
CUDA side, so parity statements that pass through these types are defined with
one implementation on both sides.
"""
from __future__ import annotations

import math
from typing import Dict, Iterator, List, Optional, Tuple

import numpy as np

try:  # pragma: no cover - exercised only where pyannote.core exists
    from pyannote.core import (  # type: ignore
        Annotation,
        Segment,
        SlidingWindow,
        SlidingWindowFeature,
    )

    HAVE_PYANNOTE_CORE = True
except Exception:  # noqa: BLE001
    HAVE_PYANNOTE_CORE = False

    class Segment:
        """Time interval ``[start, end)`` in seconds."""

        __slots__ = ("start", "end")

        def __init__(self, start: float = 0.0, end: float = 0.0):
            self.start = float(start)
            self.end = float(end)

        @property
        def duration(self) -> float:
            return self.end - self.start if self.end > self.start else 0.0

        @property
        def middle(self) -> float:
            return 0.5 * (self.start + self.end)

        def __iter__(self):
            yield self.start
            yield self.end

        def __bool__(self):
            return (self.end - self.start) > 1e-6

        def _key(self):
            return (self.start, self.end)

        def __eq__(self, other):
            return isinstance(other, Segment) and self._key() == other._key()

        def __lt__(self, other):
            return self._key() < other._key()

        def __hash__(self):
            return hash(self._key())

        def __repr__(self):
            return f"<Segment({self.start:g}, {self.end:g})>"

    class SlidingWindow:
        """Regular grid of frames: frame ``i`` covers ``[start + i*step, start + i*step + duration)``."""

        def __init__(self, duration: float = 0.030, step: float = 0.010,
                     start: float = 0.000, end: Optional[float] = None):
            if duration <= 0:
                raise ValueError("'duration' must be a float > 0.")
            if step <= 0:
                raise ValueError("'step' must be a float > 0.")
            self.duration = float(duration)
            self.step = float(step)
            self.start = float(start)
            self.end = math.inf if end is None else float(end)

        def __getitem__(self, i: int) -> Segment:
            s = self.start + i * self.step
            return Segment(s, s + self.duration)

        def closest_frame(self, t: float) -> int:
            return int(np.rint((t - self.start - 0.5 * self.duration) / self.step))

        def samples(self, from_duration: float, mode: str = "strict") -> int:
            if mode == "strict":
                return int(np.floor((from_duration - self.duration) / self.step)) + 1
            if mode == "loose":
                return int(np.floor((from_duration + self.duration) / self.step))
            if mode == "center":
                return int(np.rint(from_duration / self.step))
            raise ValueError(mode)

        def crop(self, focus: Segment, mode: str = "loose", fixed: Optional[float] = None,
                 return_ranges: bool = False):
            if mode == "loose":
                i = int(np.ceil((focus.start - self.duration - self.start) / self.step))
                if fixed is None:
                    j = int(np.floor((focus.end - self.start) / self.step))
                    rng = (i, j + 1)
                else:
                    rng = (i, i + self.samples(fixed, mode="loose"))
            elif mode == "strict":
                i = int(np.ceil((focus.start - self.start) / self.step))
                if fixed is None:
                    j = int(np.floor((focus.end - self.duration - self.start) / self.step))
                    rng = (i, j + 1)
                else:
                    rng = (i, i + self.samples(fixed, mode="strict"))
            elif mode == "center":
                i = self.closest_frame(focus.start)
                if fixed is None:
                    j = self.closest_frame(focus.end)
                    rng = (i, j + 1)
                else:
                    rng = (i, i + self.samples(fixed, mode="center"))
            else:
                raise ValueError(mode)
            if return_ranges:
                return [list(rng)]
            return np.arange(rng[0], rng[1])

        def __repr__(self):
            return f"<SlidingWindow(duration={self.duration:g}, step={self.step:g}, start={self.start:g})>"

    class SlidingWindowFeature:
        """``data`` (frames, dim) laid on a :class:`SlidingWindow`."""

        def __init__(self, data: np.ndarray, sliding_window: SlidingWindow, labels: Optional[List[str]] = None):
            self.sliding_window = sliding_window
            self.data = data
            self.labels = labels

        def __len__(self):
            return self.data.shape[0]

        def __getitem__(self, key):
            return self.data[key]

        @property
        def extent(self) -> Segment:
            # pyannote.core: SlidingWindow.range_to_segment(0, n) = [start - step/2 + duration/2, ... + n * step)
            sw = self.sliding_window
            n = self.data.shape[0]
            if sw.duration == sw.step:
                return Segment(sw.start, sw.start + n * sw.step)
            start = sw.start + (0 - 0.5) * sw.step + 0.5 * sw.duration
            return Segment(start, start + n * sw.step)

        def crop(self, focus: Segment, mode: str = "loose", fixed: Optional[float] = None,
                 return_data: bool = True):
            (lo, hi), = self.sliding_window.crop(focus, mode=mode, fixed=fixed, return_ranges=True)
            n, dim = self.data.shape[0], self.data.shape[1:]
            if fixed is None:
                lo_c, hi_c = max(0, lo), min(n, hi)
                out = self.data[lo_c:hi_c]
                first = lo_c
            else:
                # exactly hi-lo frames, edge-padded where the range leaves the data
                idx = np.clip(np.arange(lo, hi), 0, n - 1)
                out = self.data[idx]
                first = lo
            if return_data:
                return out
            sw = SlidingWindow(start=self.sliding_window[first].start,
                               duration=self.sliding_window.duration, step=self.sliding_window.step)
            return SlidingWindowFeature(out, sw, labels=self.labels)

    class Annotation:
        """``(segment, track) -> label`` store with RTTM output."""

        def __init__(self, uri: Optional[str] = None, modality: Optional[str] = None):
            self.uri = uri
            self.modality = modality
            self._tracks: Dict[Tuple[Segment, object], object] = {}

        def __setitem__(self, key, label):
            if isinstance(key, Segment):
                key = (key, "_")
            segment, track = key
            if not segment:
                return
            self._tracks[(segment, track)] = label

        def __len__(self):
            return len({s for s, _ in self._tracks})

        def itertracks(self, yield_label: bool = False) -> Iterator:
            for (segment, track) in sorted(self._tracks, key=lambda st: (st[0]._key(), str(st[1]))):
                if yield_label:
                    yield segment, track, self._tracks[(segment, track)]
                else:
                    yield segment, track

        def labels(self) -> List:
            return sorted(set(self._tracks.values()), key=str)

        def update(self, other: "Annotation", copy: bool = False) -> "Annotation":
            target = self.copy() if copy else self
            for segment, track, label in other.itertracks(yield_label=True):
                t = track
                while (segment, t) in target._tracks:
                    t = f"{t}'"
                target._tracks[(segment, t)] = label
            return target

        def copy(self) -> "Annotation":
            new = Annotation(self.uri, self.modality)
            new._tracks = dict(self._tracks)
            return new

        def support(self, collar: float = 0.0) -> "Annotation":
            """Merge same-label segments closer than ``collar`` seconds."""
            out = Annotation(self.uri, self.modality)
            by_label: Dict[object, List[Segment]] = {}
            for segment, _, label in self.itertracks(yield_label=True):
                by_label.setdefault(label, []).append(segment)
            n = 0
            for label in sorted(by_label, key=str):
                segs = sorted(by_label[label])
                cur = Segment(segs[0].start, segs[0].end)
                for s in segs[1:]:
                    if s.start - cur.end < collar or s.start <= cur.end:     # pyannote: gaps strictly shorter than the collar
                        cur = Segment(cur.start, max(cur.end, s.end))
                    else:
                        out._tracks[(cur, n)] = label
                        n += 1
                        cur = Segment(s.start, s.end)
                out._tracks[(cur, n)] = label
                n += 1
            return out

        def to_rttm(self) -> str:
            uri = self.uri if self.uri else "<NA>"
            lines = []
            for segment, _, label in self.itertracks(yield_label=True):
                lines.append(
                    f"SPEAKER {uri} 1 {segment.start:.3f} {segment.duration:.3f} <NA> <NA> {label} <NA> <NA>\n"
                )
            return "".join(lines)

        def write_rttm(self, file):
            file.write(self.to_rttm())


__all__ = ["Segment", "SlidingWindow", "SlidingWindowFeature", "Annotation", "HAVE_PYANNOTE_CORE"]


def extent_bounds(feature) -> Tuple[float, float]:
    """``(feature.extent.start, feature.extent.end)`` without building a ``Segment`` per window (the batch loops of the pipelines
    do this hundreds of times per call).  With the real ``pyannote.core`` the property itself is used."""
    if HAVE_PYANNOTE_CORE:
        e = feature.extent
        return e.start, e.end
    sw, n = feature.sliding_window, feature.data.shape[0]
    start = sw.start if sw.duration == sw.step else sw.start + (0 - 0.5) * sw.step + 0.5 * sw.duration
    return start, start + n * sw.step

