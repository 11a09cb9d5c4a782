"""RTTM sinks after the hot path (mirror the behaviour of reference ``src/diart/sinks.py:25-88``).

The reference classes are ``rx`` observers; these are plain objects with the same ``on_next`` / ``on_error`` /
``on_completed`` protocol (``rx`` is not a dependency here), so they can be subscribed to an ``rx`` pipeline as they are or
called directly on the ``(Annotation, SlidingWindowFeature)`` tuples a pipeline returns.  On-disk format:
``SPEAKER <uri> 1 <start %.3f> <duration %.3f> <NA> <NA> <label> <NA> <NA>``; same-speaker turns closer than ``patch_collar``
seconds are merged when the stream ends (``sinks.py:37-47,66-69``).
"""
from __future__ import annotations

from pathlib import Path
from typing import Dict, Optional, Tuple, Union

from .core import Annotation, Segment


def _extract_prediction(value: Union[Tuple, Annotation]) -> Annotation:
    if isinstance(value, tuple):
        return value[0]
    if isinstance(value, Annotation):
        return value
    raise ValueError(f"Expected tuple or Annotation, but got {type(value)}")


def load_rttm(path: Union[str, Path]) -> Dict[str, Annotation]:
    """RTTM file -> ``{uri: Annotation}`` (what ``pyannote.database.util.load_rttm`` returns; the reader behind the
    reference's ``RTTMWriter.patch``)"""
    out: Dict[str, Annotation] = {}
    with open(path) as f:
        for line in f:
            parts = line.split()
            if len(parts) < 8 or parts[0] != "SPEAKER":
                continue
            uri, start, duration, label = parts[1], float(parts[3]), float(parts[4]), parts[7]
            ann = out.setdefault(uri, Annotation(uri=uri, modality="speaker"))
            track = 0
            segment = Segment(start, start + duration)
            while (segment, track) in ann._tracks:
                track += 1
            ann[segment, track] = label
    return out


class _PredictionSink:
    """observer protocol shared by the sinks: every prediction gets the sink's uri and is handed to ``_add``; the end of the
    stream (completion or error) calls ``close`` once more than it has to, which is harmless"""

    uri: Optional[str] = None

    def _add(self, prediction: Annotation):
        raise NotImplementedError

    def close(self):
        raise NotImplementedError

    def on_next(self, value: Union[Tuple, Annotation]):
        prediction = _extract_prediction(value)
        prediction.uri = self.uri
        self._add(prediction)

    def on_error(self, error: Exception):
        self.close()

    def on_completed(self):
        self.close()


class PredictionAccumulator(_PredictionSink):
    """Keeps the union of all predictions in memory; ``get_prediction()`` returns it with same-speaker turns that are closer
    than ``patch_collar`` seconds merged."""

    def __init__(self, uri: Optional[str] = None, patch_collar: float = 0.05):
        self.uri = uri
        self.patch_collar = patch_collar
        self._prediction: Optional[Annotation] = None

    def _add(self, prediction: Annotation):
        if self._prediction is None:
            self._prediction = prediction
        else:
            self._prediction.update(prediction)

    def patch(self):
        if self._prediction is not None:
            self._prediction = self._prediction.support(self.patch_collar)

    close = patch

    def get_prediction(self) -> Optional[Annotation]:
        self.patch()
        return self._prediction


class RTTMWriter(_PredictionSink):
    """Appends every prediction to ``path`` as it arrives (a pre-existing file is removed first) and, when the stream ends,
    rewrites the file with close same-speaker turns merged.  The merged version is built from what the file holds, so
    lines appended by someone else in between survive, as with the reference."""

    def __init__(self, uri: str, path: Union[str, Path], patch_collar: float = 0.05):
        self.uri = uri
        self.patch_collar = patch_collar
        self.path = Path(path).expanduser()
        self.path.unlink(missing_ok=True)

    def _add(self, prediction: Annotation):
        with open(self.path, "a") as file:
            prediction.write_rttm(file)

    def patch(self):
        if not self.path.exists():
            return
        merged = PredictionAccumulator(self.uri, self.patch_collar)
        for annotation in list(load_rttm(self.path).values())[:1]:
            merged.on_next(annotation)
        prediction = merged.get_prediction()
        if prediction is not None:
            self.path.write_text(prediction.to_rttm())

    close = patch
