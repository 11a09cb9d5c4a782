"""Device-side ``rearrange_audio_stream`` (reference ``src/diart/operators.py:44-100``).

The reference turns an audio stream into 90 %-overlapping windows on the host, so every sample is stacked -- and later
uploaded -- ten times (82 MB per 256-window batch for 8.2 MB of new audio).  ``DeviceAudioStream`` keeps the stream in a
ring buffer in HBM instead: the host pushes each sample once (any block size, like the reference's sources), windows are
formed on the device, and ``SpeakerDiarization.call_stream`` / ``submit_stream`` run the hot path on them.  The host keeps
the same samples in a pinned mirror, from which the aggregated waveform of every output is sliced.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _lib


class DeviceAudioStream:
    def __init__(self, duration: float = 5, step: float = 0.5, sample_rate: int = 16000, max_windows: int = 256,
                 device: Optional[torch.device] = None, start_time: float = 0.0):
        self.sample_rate = sample_rate
        self.chunk_samples = int(round(sample_rate * duration))      # as operators.py:47-48
        self.step_samples = int(round(sample_rate * step))
        self.duration, self.step = duration, step
        self.device = torch.device(device) if device is not None else torch.device("cuda")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device() if torch.cuda.is_available() else 0)
        _lib.require_cuda(self.device)
        h = C.c_void_p()
        _lib.check(_lib.lib().dg_stream_create(self.chunk_samples, self.step_samples, int(max_windows), self.device.index,
                                               C.byref(h)))
        self._h = h
        self.start_time = float(start_time)
        self.windows_emitted = 0
        # host copy of the not-yet-dropped samples, for the aggregated waveform outputs (audio never comes back from the GPU)
        self._host = np.zeros(0, dtype=np.float32)
        self._host_first = 0                  # absolute index of self._host[0]

    def __del__(self):
        try:
            if getattr(self, "_h", None) is not None:
                _lib.lib().dg_stream_destroy(self._h)
        except Exception:  # noqa: BLE001
            pass

    @property
    def handle(self) -> C.c_void_p:
        return self._h

    def push(self, samples: np.ndarray):
        """appends a block of samples: shape (n,), (1, n) (what the reference's sources emit, operators.py:55-58) or (n, 1)"""
        x = np.asarray(samples, dtype=np.float32)
        if x.ndim == 2:
            if 1 not in x.shape:
                raise ValueError(f"Waveform must have shape (1, samples) but {x.shape} was found")
            x = x.reshape(-1)
        x = np.ascontiguousarray(x)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().dg_stream_push_host(self._h, x.ctypes.data, len(x)))
        self._host = np.concatenate([self._host, x]) if len(self._host) else x.copy()

    @property
    def available(self) -> int:
        return int(_lib.lib().dg_stream_available(self._h))

    def window_start_time(self, i: int) -> float:
        return self.start_time + i * self.step

    def host_window(self, i: int) -> np.ndarray:
        """window i of the stream from the host copy, shape (chunk_samples, 1)"""
        a = i * self.step_samples - self._host_first
        return self._host[a:a + self.chunk_samples, None]

    def advance(self, n_windows: int, keep_windows: int = 1):
        """bookkeeping after n_windows were consumed: drops host samples no later output can need"""
        self.windows_emitted += n_windows
        first_needed = max(0, self.windows_emitted - keep_windows) * self.step_samples
        if first_needed > self._host_first:
            self._host = self._host[first_needed - self._host_first:]
            self._host_first = first_needed

    def windows(self, n: int) -> torch.Tensor:
        """the next n windows as a dense (n, chunk_samples) device tensor (consumes them)"""
        out = torch.empty((n, self.chunk_samples), device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().dg_stream_windows(self._h, n, out.data_ptr(), _lib.stream_ptr(self.device)))
        self.advance(n)
        return out

    def reset(self, start_time: float = 0.0):
        _lib.check(_lib.lib().dg_stream_reset(self._h))
        self.start_time, self.windows_emitted = float(start_time), 0
        self._host, self._host_first = np.zeros(0, dtype=np.float32), 0
