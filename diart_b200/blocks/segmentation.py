"""``SpeakerSegmentation`` block (mirrors reference ``src/diart/blocks/segmentation.py:10-48``)."""
from __future__ import annotations

from typing import Optional

import torch

from ..features import TemporalFeatureFormatter, TemporalFeatures
from ..models import SegmentationModel


class SpeakerSegmentation:
    def __init__(self, model: SegmentationModel, device: Optional[torch.device] = None):
        self.model = model
        self.model.eval()
        self.device = device if device is not None else torch.device("cuda")
        self.model.to(self.device)
        self.formatter = TemporalFeatureFormatter()

    @staticmethod
    def from_pretrained(model, use_hf_token=True, device: Optional[torch.device] = None) -> "SpeakerSegmentation":
        return SpeakerSegmentation(SegmentationModel.from_pretrained(model, use_hf_token), device)

    def forward_device(self, batch: torch.Tensor) -> torch.Tensor:
        """(batch, samples) or (batch, samples, 1) on any device -> (batch, frames, speakers) on ``self.device``."""
        if batch.ndim == 3:
            batch = batch[:, :, 0]
        with torch.no_grad():
            return self.model(batch.to(self.device, non_blocking=True).unsqueeze(1))

    def __call__(self, waveform: TemporalFeatures) -> TemporalFeatures:
        """waveform (samples, channels) or (batch, samples, channels) -> (batch, frames, speakers), on the
        CPU and in the caller's feature type, exactly like the reference block."""
        wave = self.formatter.cast(waveform)                       # (batch, samples, channels)
        assert wave.shape[2] == 1, "expected mono audio"
        return self.formatter.restore_type(self.forward_device(wave).cpu())
