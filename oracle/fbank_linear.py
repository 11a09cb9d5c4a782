"""Kaldi fbank as two matrix products (TEST INFRASTRUCTURE: a CPU prototype of the B200 mapping of variant B's front end,
SURVEY.md Appendix A.6; checked against ``torchaudio.compliance.kaldi.fbank`` in ``tests/test_oracle_golden.py``).

Every step of ``kaldi.fbank`` before the power spectrum is linear in the 400 samples of a frame --
DC removal ``(I - 11^T/400)``, pre-emphasis ``x[j] - 0.97 x[max(j-1, 0)]``, the Hamming window, zero padding to 512 and
the real DFT -- so they collapse into ONE ``[514, 400]`` matrix; the frames themselves are an overlapping-row view of the
waveform (row pitch 160 samples), i.e. exactly the operand form of the shifted-window tcgen05 GEMM of this repository.
Then ``power = re^2 + im^2`` (257 bins), ``mel = melbank[80, 257] @ power`` and ``log(max(mel, eps))``.
"""
from __future__ import annotations

import math

import numpy as np


def frame_operator(window_size: int = 400, padded: int = 512, preemphasis: float = 0.97) -> np.ndarray:
    """float64 ``[2 * (padded // 2 + 1), window_size]``: rows 0..256 real part, rows 257..513 imaginary part of the DFT of
    window * preemph(frame - mean(frame)) zero-padded to 512"""
    n = window_size
    dc = np.eye(n) - np.full((n, n), 1.0 / n)
    pre = np.eye(n)
    pre[0, 0] -= preemphasis                                # x[0] - 0.97 x[0] (replicate padding)
    for j in range(1, n):
        pre[j, j - 1] -= preemphasis
    window = 0.54 - 0.46 * np.cos(2.0 * math.pi * np.arange(n) / (n - 1))        # kaldi "hamming"
    k = np.arange(padded // 2 + 1)[:, None] * np.arange(n)[None, :] * (2.0 * math.pi / padded)
    dft = np.concatenate([np.cos(k), -np.sin(k)], axis=0)
    return dft @ (window[:, None] * (pre @ dc))


def mel_banks(num_bins: int = 80, padded: int = 512, sample_freq: float = 16000.0, low_freq: float = 20.0,
              high_freq: float = 0.0) -> np.ndarray:
    """kaldi ``get_mel_banks`` without VTLN: triangular filters on the mel scale 1127 ln(1 + f / 700); ``[num_bins, padded/2 + 1]``
    (the last column, the Nyquist bin, is zero as in torchaudio's right-padding)"""
    nyquist = 0.5 * sample_freq
    if high_freq <= 0.0:
        high_freq += nyquist
    mel = lambda f: 1127.0 * np.log(1.0 + f / 700.0)
    num_fft_bins = padded // 2
    fft_bin_width = sample_freq / padded
    mel_low, mel_high = mel(low_freq), mel(high_freq)
    delta = (mel_high - mel_low) / (num_bins + 1)
    b = np.arange(num_bins)[:, None]
    left, center, right = mel_low + b * delta, mel_low + (b + 1.0) * delta, mel_low + (b + 2.0) * delta
    m = mel(fft_bin_width * np.arange(num_fft_bins))[None, :]
    up, down = (m - left) / (center - left), (right - m) / (right - center)
    banks = np.maximum(0.0, np.minimum(up, down))
    return np.pad(banks, ((0, 0), (0, 1)))


def fbank(waveform: np.ndarray, window_size: int = 400, shift: int = 160) -> np.ndarray:
    """waveform (S,) already scaled to int16 range -> (frames, 80) log mel energies, float32 arithmetic like the reference"""
    x = np.asarray(waveform, dtype=np.float32)
    frames = 1 + (x.shape[0] - window_size) // shift
    rows = np.lib.stride_tricks.as_strided(x, (frames, window_size), (shift * x.strides[0], x.strides[0]))   # overlapping rows
    spec = rows @ frame_operator(window_size).astype(np.float32).T
    half = spec.shape[1] // 2
    power = spec[:, :half] ** 2 + spec[:, half:] ** 2
    mel = power @ mel_banks().astype(np.float32).T
    return np.log(np.maximum(mel, np.finfo(np.float32).eps))
