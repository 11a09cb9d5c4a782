"""``VoiceActivityDetection`` pipeline (mirrors reference ``src/diart/blocks/vad.py:27-191``): the segmentation
network of the hot path followed by a max over local speakers, aggregation and binarisation (SURVEY.md 8(f) row 4:
it falls out of the native segmentation block).  Output label of every speech turn is ``"speech"``."""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from .. import models as m
from ..core import Annotation, Segment, SlidingWindow, SlidingWindowFeature
from . import base
from .aggregation import DelayedAggregation
from .segmentation import SpeakerSegmentation
from .utils import Binarize


class VoiceActivityDetectionConfig(base.PipelineConfig):
    def __init__(self, segmentation: Optional[m.SegmentationModel] = None, duration: float = 5, step: float = 0.5,
                 latency=None, tau_active: float = 0.6, device: Optional[torch.device] = None,
                 sample_rate: int = 16000, **kwargs):
        self.segmentation = segmentation or m.SegmentationModel.from_pyannote("pyannote/segmentation")
        self._duration, self._step, self._sample_rate = duration, step, sample_rate
        self._latency = latency
        if self._latency is None or self._latency == "min":
            self._latency = self._step
        elif self._latency == "max":
            self._latency = self._duration
        self.tau_active = tau_active
        self.device = device or torch.device("cuda")

    @property
    def duration(self) -> float:
        return self._duration

    @property
    def step(self) -> float:
        return self._step

    @property
    def latency(self) -> float:
        return self._latency

    @property
    def sample_rate(self) -> int:
        return self._sample_rate


class VoiceActivityDetection(base.Pipeline):
    def __init__(self, config: Optional[VoiceActivityDetectionConfig] = None):
        self._config = VoiceActivityDetectionConfig() if config is None else config
        msg = f"Latency should be in the range [{self._config.step}, {self._config.duration}]"
        assert self._config.step <= self._config.latency <= self._config.duration, msg
        self.segmentation = SpeakerSegmentation(self._config.segmentation, self._config.device)
        self.pred_aggregation = DelayedAggregation(self._config.step, self._config.latency, strategy="hamming",
                                                   cropping_mode="loose")
        self.audio_aggregation = DelayedAggregation(self._config.step, self._config.latency, strategy="first",
                                                    cropping_mode="center")
        self.binarize = Binarize(self._config.tau_active)
        self.timestamp_shift = 0
        self.chunk_buffer, self.pred_buffer = [], []

    @staticmethod
    def get_config_class() -> type:
        return VoiceActivityDetectionConfig

    @staticmethod
    def suggest_metric():
        from pyannote.metrics.detection import DetectionErrorRate  # optional dependency

        return DetectionErrorRate(collar=0, skip_overlap=False)

    @staticmethod
    def hyper_parameters() -> Sequence[base.HyperParameter]:
        return [base.TauActive]

    @property
    def config(self) -> VoiceActivityDetectionConfig:
        return self._config

    def reset(self):
        self.set_timestamp_shift(0)
        self.chunk_buffer, self.pred_buffer = [], []

    def set_timestamp_shift(self, shift: float):
        self.timestamp_shift = shift

    def __call__(self, waveforms: Sequence[SlidingWindowFeature]) -> Sequence[Tuple[Annotation, SlidingWindowFeature]]:
        assert len(waveforms) >= 1, "Pipeline expected at least 1 input"
        batch = torch.stack([torch.from_numpy(np.asarray(w.data, dtype=np.float32)) for w in waveforms])
        expected = int(np.rint(self.config.duration * self.config.sample_rate))
        assert batch.shape[1] == expected, f"Expected {expected} samples per chunk, but got {batch.shape[1]}"
        scores = self.segmentation.forward_device(batch)                     # (batch, frames, speakers) on the device
        vads = torch.max(scores, dim=-1, keepdim=True)[0].cpu().numpy()      # (batch, frames, 1)
        resolution = waveforms[0].extent.duration / vads.shape[1]
        outputs = []
        for wav, vad in zip(waveforms, vads):
            sw = SlidingWindow(start=wav.extent.start, duration=resolution, step=resolution)
            self.chunk_buffer.append(wav)
            self.pred_buffer.append(SlidingWindowFeature(vad, sw))
            agg_waveform = self.audio_aggregation(self.chunk_buffer)
            turns = self.binarize(self.pred_aggregation(self.pred_buffer))
            speech = Annotation(uri=turns.uri, modality="speech")
            for n, (segment, _) in enumerate(turns.itertracks()):
                shifted = Segment(segment.start + self.timestamp_shift, segment.end + self.timestamp_shift)
                speech[shifted, n] = "speech"
            outputs.append((speech, agg_waveform))
            if len(self.chunk_buffer) == self.pred_aggregation.num_overlapping_windows:
                self.chunk_buffer = self.chunk_buffer[1:]
                self.pred_buffer = self.pred_buffer[1:]
        return outputs
