"""Embedding-side blocks (mirror reference ``src/diart/blocks/embedding.py``):
``SpeakerEmbedding``, ``OverlappedSpeechPenalty``, ``EmbeddingNormalization`` and the composite
``OverlapAwareSpeakerEmbedding``.  With a ``B200XVectorSincNet`` behind the model the composite runs
fused on the device (one trunk pass per chunk instead of one per local speaker)."""
from __future__ import annotations

from typing import Optional, Union

import torch

from .. import _lib
from ..features import TemporalFeatureFormatter, TemporalFeatures
from ..models import B200XVectorSincNet, EmbeddingModel


def _device_of(device: Optional[torch.device]) -> torch.device:
    return device if device is not None else torch.device("cuda")


class SpeakerEmbedding:
    def __init__(self, model: EmbeddingModel, device: Optional[torch.device] = None):
        self.model = model
        self.model.eval()
        self.device = _device_of(device)
        self.model.to(self.device)
        self.waveform_formatter = TemporalFeatureFormatter()
        self.weights_formatter = TemporalFeatureFormatter()

    @staticmethod
    def from_pretrained(model, use_hf_token=True, device: Optional[torch.device] = None) -> "SpeakerEmbedding":
        return SpeakerEmbedding(EmbeddingModel.from_pretrained(model, use_hf_token), device)

    @property
    def native(self) -> Optional[B200XVectorSincNet]:
        inner = getattr(self.model, "model", None)
        return inner if isinstance(inner, B200XVectorSincNet) else None

    def forward_device(self, batch: torch.Tensor, weights: Optional[torch.Tensor], normalize: bool = False,
                       norm: float = 1.0) -> torch.Tensor:
        """batch (B,S) / (B,S,1), weights (B,F,K) or None -> (B,K,D) (or (B,D)) on ``self.device``."""
        if batch.ndim == 3:
            batch = batch[:, :, 0]
        x = batch.to(self.device, non_blocking=True)
        with torch.no_grad():
            if weights is None:
                out = self.model(x.unsqueeze(1))
            elif self.native is not None:
                return self.native.forward_fused(x, weights.to(self.device), normalize, norm)
            else:  # any other EmbeddingModel: the reference's K-fold repeat (embedding.py:57-59)
                B, F, K = weights.shape
                w = weights.to(self.device).permute(0, 2, 1).reshape(B * K, F)
                out = self.model(x.repeat_interleave(K, dim=0).unsqueeze(1), w).reshape(B, K, -1)
            if normalize:
                out = norm * out / torch.norm(out, p=2, dim=-1, keepdim=True)
            return out

    def __call__(self, waveform: TemporalFeatures, weights: Optional[TemporalFeatures] = None) -> torch.Tensor:
        wave = self.waveform_formatter.cast(waveform)
        w = None if weights is None else self.weights_formatter.cast(weights)
        return self.forward_device(wave, w).squeeze().cpu()


class OverlappedSpeechPenalty:
    """Eq. 2 of the paper: ``w = s^gamma * softmax(beta * s)^gamma`` clamped at 1e-8, optional min-max
    normalisation over frames (reference ``embedding.py:71-107``, ``functional.py:6-13``).  Computed by
    ``dg_osp`` on the device the input lives on (CPU inputs are moved to ``device`` and back)."""

    def __init__(self, gamma: float = 3, beta: float = 10, normalize: bool = False,
                 device: Optional[torch.device] = None):
        self.gamma, self.beta, self.normalize = gamma, beta, normalize
        self.device = _device_of(device)
        self.formatter = TemporalFeatureFormatter()

    def forward_device(self, seg: torch.Tensor) -> torch.Tensor:
        _lib.require_cuda(seg.device)
        seg = seg.contiguous()
        B, F, K = seg.shape
        out = torch.empty_like(seg)
        with torch.cuda.device(seg.device):
            _lib.check(_lib.lib().dg_osp(seg.data_ptr(), B, F, K, float(self.gamma), float(self.beta),
                                         int(self.normalize), out.data_ptr(), _lib.stream_ptr(seg.device)))
        return out

    def __call__(self, segmentation: TemporalFeatures) -> TemporalFeatures:
        seg = self.formatter.cast(segmentation)
        src = seg.device
        out = self.forward_device(seg.to(self.device) if src.type != "cuda" else seg)
        return self.formatter.restore_type(out.to(src))


class EmbeddingNormalization:
    """``norm * e / ||e||`` (reference ``embedding.py:110-120``, ``functional.py:16-27``)."""

    def __init__(self, norm: Union[float, torch.Tensor] = 1, device: Optional[torch.device] = None):
        self.norm = norm
        if isinstance(self.norm, torch.Tensor) and self.norm.ndim == 2:
            self.norm = self.norm.unsqueeze(0)
        self.device = _device_of(device)

    def __call__(self, embeddings: torch.Tensor) -> torch.Tensor:
        if embeddings.ndim == 2:
            embeddings = embeddings.unsqueeze(0)
        if isinstance(self.norm, torch.Tensor):
            b1, s1, _ = self.norm.shape
            b2, s2, _ = embeddings.shape
            assert b1 == b2 and s1 == s2
            e = embeddings.to(self.device)
            return (self.norm.to(self.device) * e / torch.norm(e, p=2, dim=-1, keepdim=True)).to(embeddings.device)
        src = embeddings.device
        e = embeddings.to(self.device, torch.float32).contiguous()
        out = torch.empty_like(e)
        rows, D = e.shape[0] * e.shape[1], e.shape[2]
        with torch.cuda.device(e.device):
            _lib.check(_lib.lib().dg_normalize_embeddings(e.data_ptr(), rows, D, float(self.norm), out.data_ptr(),
                                                          _lib.stream_ptr(e.device)))
        return out.to(src)


class OverlapAwareSpeakerEmbedding:
    def __init__(self, model: EmbeddingModel, gamma: float = 3, beta: float = 10,
                 norm: Union[float, torch.Tensor] = 1, normalize_weights: bool = False,
                 device: Optional[torch.device] = None):
        self.embedding = SpeakerEmbedding(model, device)
        self.osp = OverlappedSpeechPenalty(gamma, beta, normalize_weights, device)
        self.normalize = EmbeddingNormalization(norm, device)

    @staticmethod
    def from_pretrained(model, gamma: float = 3, beta: float = 10, norm: Union[float, torch.Tensor] = 1,
                        use_hf_token=True, normalize_weights: bool = False, device: Optional[torch.device] = None):
        model = EmbeddingModel.from_pretrained(model, use_hf_token)
        return OverlapAwareSpeakerEmbedding(model, gamma, beta, norm, normalize_weights, device)

    def forward_device(self, batch: torch.Tensor, segmentation: torch.Tensor) -> torch.Tensor:
        """batch (B,S[,1]) and segmentation (B,F,K) on the device -> normalised embeddings (B,K,D) on the device."""
        weights = self.osp.forward_device(segmentation.to(self.embedding.device))
        scalar = not isinstance(self.normalize.norm, torch.Tensor)
        out = self.embedding.forward_device(batch, weights, normalize=scalar,
                                            norm=float(self.normalize.norm) if scalar else 1.0)
        return out if scalar else self.normalize(out)

    def __call__(self, waveform: TemporalFeatures, segmentation: TemporalFeatures) -> torch.Tensor:
        wave = self.embedding.waveform_formatter.cast(waveform)
        seg = self.embedding.weights_formatter.cast(segmentation)
        return self.forward_device(wave, seg).cpu()
