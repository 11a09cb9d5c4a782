"""Pipeline contract (mirrors reference ``src/diart/blocks/base.py:13-137``): what
``StreamingInference`` / ``Benchmark`` call on a pipeline and its config."""
from __future__ import annotations

from abc import ABC, abstractmethod
from dataclasses import dataclass
from typing import Any, Sequence, Tuple

from ..core import SlidingWindowFeature


@dataclass
class HyperParameter:
    name: str
    low: float
    high: float

    @staticmethod
    def from_name(name: str) -> "HyperParameter":
        for hp in (TauActive, RhoUpdate, DeltaNew):
            if hp.name == name:
                return hp
        raise ValueError(f"Hyper-parameter '{name}' not recognized")


TauActive = HyperParameter("tau_active", low=0, high=1)
RhoUpdate = HyperParameter("rho_update", low=0, high=1)
DeltaNew = HyperParameter("delta_new", low=0, high=2)


class PipelineConfig(ABC):
    @property
    @abstractmethod
    def duration(self) -> float: ...

    @property
    @abstractmethod
    def step(self) -> float: ...

    @property
    @abstractmethod
    def latency(self) -> float: ...

    @property
    @abstractmethod
    def sample_rate(self) -> int: ...

    def get_file_padding(self, filepath=None, file_duration: float = None) -> Tuple[float, float]:
        """(left, right) padding in seconds so that a file yields whole chunks
        (reference ``base.py:81-85`` with ``utils.get_padding_left/right``)."""
        if file_duration is None:
            import torchaudio  # only needed when padding is derived from a file on disk

            info = torchaudio.load(str(filepath))
            file_duration = info[0].shape[1] / info[1]
        right = self.latency - self.step
        total = file_duration + right
        left = max(0.0, self.duration - total)
        return left, right


class WindowTiming(PipelineConfig):
    """Chunk geometry shared by the pipeline configurations: ``duration`` / ``step`` in seconds, ``latency`` in seconds or
    ``"min"`` (= step, the default) / ``"max"`` (= duration), as in the reference's config constructors
    (``blocks/diarization.py:33-60``, ``blocks/vad.py:27-65``)."""

    def _set_timing(self, duration: float, step: float, latency, sample_rate: int):
        self._duration, self._step, self._sample_rate = duration, step, sample_rate
        self._latency = step if latency in (None, "min") else (duration if latency == "max" else latency)

    duration = property(lambda self: self._duration)
    step = property(lambda self: self._step)
    latency = property(lambda self: self._latency)
    sample_rate = property(lambda self: self._sample_rate)


class Pipeline(ABC):
    @staticmethod
    @abstractmethod
    def get_config_class() -> type: ...

    @staticmethod
    @abstractmethod
    def suggest_metric(): ...

    @staticmethod
    @abstractmethod
    def hyper_parameters() -> Sequence[HyperParameter]: ...

    @property
    @abstractmethod
    def config(self) -> PipelineConfig: ...

    @abstractmethod
    def reset(self): ...

    @abstractmethod
    def set_timestamp_shift(self, shift: float): ...

    @abstractmethod
    def __call__(self, waveforms: Sequence[SlidingWindowFeature]) -> Sequence[Tuple[Any, SlidingWindowFeature]]: ...
