"""Model-loader plugin boundary (mirrors reference ``src/diart/models.py``).

diart's blocks accept any ``SegmentationModel(loader)`` / ``EmbeddingModel(loader)`` where ``loader``
is a ``Callable[[], Callable]`` (reference ``models.py:112-133``, documented plugin API
``README.md:186-209``).  This module provides

* ``LazyModel`` / ``SegmentationModel`` / ``EmbeddingModel`` with the reference's interface
  (lazy ``load``, ``to``, ``eval``, ``__call__``), so the package is usable where diart is not
  installed, and
* ``B200PyanNet`` / ``B200XVectorSincNet``: the loaded callables, backed by libdiartb200.so, plus
  loader classes that build them from a pyannote ``state_dict`` (dict, ``.pt`` / ``.npz`` file, or a
  pyannote checkpoint when ``pyannote.audio`` is installed).  The same loaders can be handed to the
  *unmodified* reference classes: ``diart.models.SegmentationModel(B200SegmentationLoader(sd))``.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path
from typing import Callable, Dict, Optional, Tuple, Union

import numpy as np
import torch

from . import _lib

StateDict = Dict[str, Union[torch.Tensor, np.ndarray]]


def _load_pyannote(name: str, use_hf_token=True):
    """pyannote model by name (what reference ``src/diart/models.py:41-53`` does), with the token the caller gave"""
    try:
        from pyannote.audio import Model  # type: ignore
    except ImportError as e:
        raise FileNotFoundError(
            f"'{name}' is not a state-dict file and pyannote.audio is not installed to fetch it") from e
    return Model.from_pretrained(name, use_auth_token=use_hf_token)


def _load_state(source, use_hf_token=True):
    """-> (state dict, the pyannote model it came from or None)"""
    if isinstance(source, dict):
        return source, None
    if isinstance(source, (str, Path)):
        path = Path(source)
        if path.suffix == ".npz":
            return dict(np.load(path)), None
        if path.exists():
            obj = torch.load(path, map_location="cpu", weights_only=True)
            return (obj.get("state_dict", obj) if isinstance(obj, dict) else obj), None
        model = _load_pyannote(str(source), use_hf_token)      # a pyannote model name, e.g. "pyannote/segmentation"
        return model.state_dict(), model
    if hasattr(source, "state_dict"):
        return source.state_dict(), source
    raise ValueError("expected a state dict, a path to one, a pyannote model name or an nn.Module")


def _powerset_classes(num_speakers: int, max_per_frame: int) -> int:
    from math import comb

    return sum(comb(num_speakers, k) for k in range(max_per_frame + 1))


class _Handle:
    """Owns a libdiartb200 handle; created on the first ``to(cuda_device)``."""

    def __init__(self, state: StateDict):
        self._state = state
        self._h: Optional[C.c_void_p] = None
        self.device: Optional[torch.device] = None

    def _create(self, device: torch.device) -> C.c_void_p:
        raise NotImplementedError

    def _destroy(self, h):
        raise NotImplementedError

    def to(self, device: Union[str, torch.device]):
        device = torch.device(device)
        _lib.require_cuda(device)
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        if self._h is not None and self.device == device:
            return self
        if self._h is not None:
            self._destroy(self._h)
            self._h = None
        self._h = self._create(device)
        self.device = device
        return self

    def eval(self):
        return self

    @property
    def handle(self) -> C.c_void_p:
        if self._h is None:
            raise _lib.DiartB200Error("model is not on a CUDA device yet: call .to(torch.device('cuda')) first")
        return self._h

    def __del__(self):
        try:
            if self._h is not None:
                self._destroy(self._h)
        except Exception:  # noqa: BLE001 - interpreter shutdown
            pass


class B200PyanNet(_Handle):
    """pyannote/segmentation forward on the GPU: ``(B, 1, S) cuda float32 -> (B, F, K)``.

    ``powerset=(num_speakers, max_speakers_per_frame)`` declares a powerset model (pyannote/segmentation-3.0): the
    classifier's outputs are subsets of the local speakers and the forward returns the hard multilabel scores the
    reference's ``PowersetAdapter`` produces (reference ``src/diart/models.py:29-39``)."""

    def __init__(self, state: StateDict, powerset: Optional[Tuple[int, int]] = None):
        super().__init__(state)
        self.powerset = powerset

    def _create(self, device):
        arr, n, keep = _lib.pack_state_dict(self._state)
        h = C.c_void_p()
        _lib.check(_lib.lib().dg_seg_create(arr, n, device.index, C.byref(h)))
        if self.powerset is not None:
            try:
                _lib.check(_lib.lib().dg_seg_set_powerset(h, int(self.powerset[0]), int(self.powerset[1])))
            except Exception:
                _lib.lib().dg_seg_destroy(h)
                raise
        return h

    def _destroy(self, h):
        _lib.lib().dg_seg_destroy(h)

    def dims(self, num_samples: int):
        f, k = C.c_int(), C.c_int()
        _lib.check(_lib.lib().dg_seg_dims(self.handle, num_samples, C.byref(f), C.byref(k)))
        return f.value, k.value

    def __call__(self, waveform: torch.Tensor) -> torch.Tensor:
        if waveform.ndim == 3:
            assert waveform.shape[1] == 1, "expected mono audio of shape (batch, 1, samples)"
            waveform = waveform[:, 0, :]
        assert waveform.ndim == 2, "expected waveform of shape (batch, 1, samples)"
        x = waveform.to(self.device, torch.float32).contiguous()
        B, S = x.shape
        F, K = self.dims(S)
        out = torch.empty((B, F, K), device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().dg_seg_forward(self.handle, x.data_ptr(), B, S, out.data_ptr(),
                                                 _lib.stream_ptr(self.device)))
        return out


class B200XVectorSincNet(_Handle):
    """pyannote/embedding forward on the GPU.

    ``__call__(waveform (N,1,S), weights (N,F) | None) -> (N,D)`` is the reference loader contract
    (``models.py:248-265``); ``forward_fused(waveform (B,S), weights (B,F,K)) -> (B,K,D)`` runs the
    trunk once per waveform.
    """

    def __init__(self, state: StateDict, pool_mode: str = "3.1"):
        super().__init__(state)
        assert pool_mode in ("3.1", "2.1")
        self.pool_mode = pool_mode

    def _create(self, device):
        arr, n, keep = _lib.pack_state_dict(self._state)
        h = C.c_void_p()
        _lib.check(_lib.lib().dg_emb_create(arr, n, 31 if self.pool_mode == "3.1" else 21, device.index, C.byref(h)))
        return h

    def _destroy(self, h):
        _lib.lib().dg_emb_destroy(h)

    def dims(self, num_samples: int):
        f, d = C.c_int(), C.c_int()
        _lib.check(_lib.lib().dg_emb_dims(self.handle, num_samples, C.byref(f), C.byref(d)))
        return f.value, d.value

    @staticmethod
    def _wave2d(waveform: torch.Tensor) -> torch.Tensor:
        if waveform.ndim == 3:
            assert waveform.shape[1] == 1, "expected mono audio of shape (batch, 1, samples)"
            waveform = waveform[:, 0, :]
        assert waveform.ndim == 2
        return waveform

    def __call__(self, waveform: torch.Tensor, weights: Optional[torch.Tensor] = None) -> torch.Tensor:
        x = self._wave2d(waveform).to(self.device, torch.float32).contiguous()
        N, S = x.shape
        _, D = self.dims(S)
        w, F = None, 0
        if weights is not None:
            w = weights.to(self.device, torch.float32).contiguous()
            assert w.ndim == 2 and w.shape[0] == N, "weights must have shape (batch, frames)"
            F = w.shape[1]
        out = torch.empty((N, D), device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().dg_emb_forward_rows(self.handle, x.data_ptr(), _lib.ptr(w), N, S, F,
                                                      out.data_ptr(), _lib.stream_ptr(self.device)))
        return out

    def forward_fused(self, waveform: torch.Tensor, weights: torch.Tensor, normalize: bool = False,
                      norm: float = 1.0) -> torch.Tensor:
        x = self._wave2d(waveform).to(self.device, torch.float32).contiguous()
        w = weights.to(self.device, torch.float32).contiguous()
        B, S = x.shape
        assert w.ndim == 3 and w.shape[0] == B, "weights must have shape (batch, frames, speakers)"
        _, F, K = w.shape
        _, D = self.dims(S)
        out = torch.empty((B, K, D), device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().dg_emb_forward(self.handle, x.data_ptr(), w.data_ptr(), B, S, F, K,
                                                 int(normalize), float(norm), out.data_ptr(),
                                                 _lib.stream_ptr(self.device)))
        return out


def _powerset_of(source) -> Optional[Tuple[int, int]]:
    """(num_speakers, max_speakers_per_frame) when ``source`` is a pyannote model with powerset specifications (what the
    reference's loader checks, ``src/diart/models.py:50-53``), else None"""
    specs = getattr(source, "specifications", None)
    if specs is not None and getattr(specs, "powerset", False):
        return len(specs.classes), int(specs.powerset_max_classes)
    return None


class B200SegmentationLoader:
    """``loader`` argument for ``SegmentationModel`` (ours or the reference's).  ``powerset=(num_speakers,
    max_speakers_per_frame)`` for powerset checkpoints given as plain state dicts (it is read from the model's
    specifications when ``source`` is a pyannote model)."""

    def __init__(self, source, powerset: Optional[Tuple[int, int]] = None, use_hf_token=True):
        self.source = source
        self.powerset = powerset
        self.use_hf_token = use_hf_token

    def __call__(self) -> B200PyanNet:
        # the model is loaded ONCE; its specifications (powerset or multilabel) are read from the loaded object before the
        # state dict is taken -- a powerset checkpoint given by name must not be decoded as 7 "speakers" through the sigmoid head
        state, model = _load_state(self.source, self.use_hf_token)
        powerset = self.powerset if self.powerset is not None else _powerset_of(model if model is not None else self.source)
        bias = state.get("classifier.bias")
        if bias is not None and powerset is not None:
            width, want = int(np.prod(tuple(bias.shape))), _powerset_classes(*powerset)
            if width != want:
                raise ValueError(f"classifier has {width} outputs but a powerset of {powerset[0]} speakers with at most "
                                 f"{powerset[1]} per frame has {want} classes")
        return B200PyanNet(state, powerset=powerset)


class B200EmbeddingLoader:
    def __init__(self, source, pool_mode: str = "3.1", use_hf_token=True):
        self.source, self.pool_mode, self.use_hf_token = source, pool_mode, use_hf_token

    def __call__(self) -> B200XVectorSincNet:
        """pyannote/embedding (XVectorSincNet) and pyannote/wespeaker-voxceleb-resnet34-LM (WeSpeakerResNet34, variant B) are
        told apart by the key names of the state dict (``dg_emb_create``)."""
        return B200XVectorSincNet(_load_state(self.source, self.use_hf_token)[0], self.pool_mode)


class LazyModel:
    """reference ``models.py:112-139``: loads on first use, forwards ``to`` / ``eval`` / ``__call__``."""

    def __init__(self, loader: Callable[[], Callable]):
        self.get_model = loader
        self.model: Optional[Callable] = None

    def is_in_memory(self) -> bool:
        return self.model is not None

    def load(self):
        if not self.is_in_memory():
            self.model = self.get_model()

    def to(self, device: torch.device) -> "LazyModel":
        self.load()
        self.model = self.model.to(device)
        return self

    def __call__(self, *args, **kwargs):
        self.load()
        return self.model(*args, **kwargs)

    def eval(self) -> "LazyModel":
        self.load()
        if isinstance(self.model, torch.nn.Module):
            self.model.eval()
        return self


class SegmentationModel(LazyModel):
    """reference ``models.py:142-198``.  ``from_pretrained`` accepts a state-dict file / dict, or a
    pyannote model name when ``pyannote.audio`` is installed (weights are then re-hosted on the GPU path)."""

    @staticmethod
    def from_pyannote(model, use_hf_token=True) -> "SegmentationModel":
        return SegmentationModel(B200SegmentationLoader(model, use_hf_token=use_hf_token))

    @staticmethod
    def from_pretrained(model, use_hf_token=True) -> "SegmentationModel":
        return SegmentationModel(B200SegmentationLoader(model, use_hf_token=use_hf_token))

    def __call__(self, waveform: torch.Tensor) -> torch.Tensor:
        return super().__call__(waveform)


class EmbeddingModel(LazyModel):
    """reference ``models.py:201-265``."""

    @staticmethod
    def from_pyannote(model, use_hf_token=True, pool_mode: str = "3.1") -> "EmbeddingModel":
        return EmbeddingModel(B200EmbeddingLoader(model, pool_mode, use_hf_token=use_hf_token))

    @staticmethod
    def from_pretrained(model, use_hf_token=True, pool_mode: str = "3.1") -> "EmbeddingModel":
        return EmbeddingModel(B200EmbeddingLoader(model, pool_mode, use_hf_token=use_hf_token))

    def __call__(self, waveform: torch.Tensor, weights: Optional[torch.Tensor] = None) -> torch.Tensor:
        embeddings = super().__call__(waveform, weights)
        if isinstance(embeddings, np.ndarray):
            embeddings = torch.from_numpy(embeddings)
        return embeddings
