"""Seeded synthetic inputs: random-init weights in pyannote ``state_dict`` layout and a speech-like
audio stream.  No checkpoints or datasets exist offline (the HF models are
gated, reference ``README.md:101-109``), so tests and ``bench.py`` run on these.

The weights are plain uniform(+-1/sqrt(fan_in)) draws, rescaled / calibrated (``synth_calib.npz``,
produced by ``oracle/calibrate.py``) so that the untrained networks produce segmentation scores that
straddle the clustering thresholds and embeddings whose cosine distances spread over (0, 2);
otherwise every sigmoid idles at 0.5 and no clustering branch is exercised.
"""
from __future__ import annotations

import math
import os
from typing import Dict

import numpy as np
import torch

_CALIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "synth_calib.npz")


def _uniform(g: torch.Generator, shape, fan_in: int, scale: float = 1.0) -> torch.Tensor:
    bound = scale / math.sqrt(fan_in)
    return (torch.rand(shape, generator=g) * 2 - 1) * bound


def _sincnet_state(g: torch.Generator, prefix: str, wn=(1.0, 0.0)) -> Dict[str, torch.Tensor]:
    to_mel = lambda hz: 2595 * np.log10(1 + hz / 700)
    to_hz = lambda mel: 700 * (10 ** (mel / 2595) - 1)
    mel = np.linspace(to_mel(30), to_mel(16000 / 2 - 100), 41, dtype="float32")
    hz = to_hz(mel)
    s = {
        prefix + "wav_norm1d.weight": torch.tensor([wn[0]]),
        prefix + "wav_norm1d.bias": torch.tensor([wn[1]]),
        prefix + "conv1d.0.filterbank.low_hz_": torch.from_numpy(hz[:-1].astype(np.float32)).view(-1, 1).clone(),
        prefix + "conv1d.0.filterbank.band_hz_": torch.from_numpy(np.diff(hz).astype(np.float32)).view(-1, 1).clone(),
        prefix + "conv1d.1.weight": _uniform(g, (60, 80, 5), 400),
        prefix + "conv1d.1.bias": _uniform(g, (60,), 400),
        prefix + "conv1d.2.weight": _uniform(g, (60, 60, 5), 300),
        prefix + "conv1d.2.bias": _uniform(g, (60,), 300),
    }
    for i, c in enumerate((80, 60, 60)):
        s[prefix + f"norm1d.{i}.weight"] = 1.0 + 0.1 * torch.randn(c, generator=g)
        s[prefix + f"norm1d.{i}.bias"] = 0.1 * torch.randn(c, generator=g)
    return s


def segmentation_state(seed: int = 4321, num_speakers: int = 3, calibrated: bool = True) -> Dict[str, torch.Tensor]:
    """state_dict of pyannote's PyanNet (pyannote/segmentation hyper-parameters), 1 472 749 parameters."""
    g = torch.Generator().manual_seed(seed)
    s = _sincnet_state(g, "sincnet.", wn=(1.25, 0.05))
    for layer in range(4):
        n_in = 60 if layer == 0 else 256
        for sfx in ("", "_reverse"):
            # x3: an untrained LSTM with default-scale weights is nearly input-independent
            s[f"lstm.weight_ih_l{layer}{sfx}"] = _uniform(g, (512, n_in), 128, 3.0)
            s[f"lstm.weight_hh_l{layer}{sfx}"] = _uniform(g, (512, 128), 128, 3.0)
            s[f"lstm.bias_ih_l{layer}{sfx}"] = _uniform(g, (512,), 128)
            s[f"lstm.bias_hh_l{layer}{sfx}"] = _uniform(g, (512,), 128)
    s["linear.0.weight"] = _uniform(g, (128, 256), 256, 2.0)
    s["linear.0.bias"] = _uniform(g, (128,), 256)
    s["linear.1.weight"] = _uniform(g, (128, 128), 128, 2.0)
    s["linear.1.bias"] = _uniform(g, (128,), 128)
    s["classifier.weight"] = _uniform(g, (num_speakers, 128), 128)
    s["classifier.bias"] = _uniform(g, (num_speakers,), 128)
    if calibrated and os.path.exists(_CALIB):
        c = np.load(_CALIB)
        if f"seg{seed}_scale" in c.files and c[f"seg{seed}_scale"].shape[0] == num_speakers:
            scale = torch.from_numpy(c[f"seg{seed}_scale"]).float()
            s["classifier.weight"] = s["classifier.weight"] * scale[:, None]
            s["classifier.bias"] = torch.from_numpy(c[f"seg{seed}_bias"]).float()
    return s


TDNN = [(60, 512, 5, 1), (512, 512, 3, 2), (512, 512, 3, 3), (512, 512, 1, 1), (512, 1500, 1, 1)]


def embedding_state(seed: int = 8765, dimension: int = 512, calibrated: bool = True) -> Dict[str, torch.Tensor]:
    """state_dict of pyannote's XVectorSincNet (pyannote/embedding), 4 346 366 parameters."""
    g = torch.Generator().manual_seed(seed)
    s = _sincnet_state(g, "sincnet.", wn=(0.8, -0.02))
    for i, (cin, cout, k, _) in enumerate(TDNN):
        s[f"tdnns.{3 * i}.weight"] = _uniform(g, (cout, cin, k), cin * k)
        s[f"tdnns.{3 * i}.bias"] = _uniform(g, (cout,), cin * k)
        s[f"tdnns.{3 * i + 2}.weight"] = 1.0 + 0.2 * torch.randn(cout, generator=g)
        s[f"tdnns.{3 * i + 2}.bias"] = 0.1 * torch.randn(cout, generator=g)
        s[f"tdnns.{3 * i + 2}.running_mean"] = 0.1 * torch.randn(cout, generator=g)
        s[f"tdnns.{3 * i + 2}.running_var"] = 0.5 + torch.rand(cout, generator=g)
    s["embedding.weight"] = _uniform(g, (dimension, 3000), 3000)
    s["embedding.bias"] = _uniform(g, (dimension,), 3000)
    if calibrated and os.path.exists(_CALIB):
        c = np.load(_CALIB)
        if f"emb{seed}_bias" in c.files and c[f"emb{seed}_bias"].shape[0] == dimension:
            for i in range(5):
                s[f"tdnns.{3 * i + 2}.running_mean"] = torch.from_numpy(c[f"emb{seed}_bn{i}_mean"]).float()
                s[f"tdnns.{3 * i + 2}.running_var"] = torch.from_numpy(c[f"emb{seed}_bn{i}_var"]).float()
            s["embedding.bias"] = torch.from_numpy(c[f"emb{seed}_bias"]).float()
    return s


def wespeaker_state(seed: int = 2468, dimension: int = 256) -> Dict[str, torch.Tensor]:
    """state_dict of pyannote's WeSpeakerResNet34 (pyannote/wespeaker-voxceleb-resnet34-LM layout, 6 634 336 parameters):
    seeded uniform(+-1/sqrt(fan_in)) convolutions, randomised BatchNorm statistics (so that the eval-mode affine is not
    the identity)."""
    g = torch.Generator().manual_seed(seed)
    s: Dict[str, torch.Tensor] = {}

    def bn(prefix: str, c: int):
        s[prefix + ".weight"] = 1.0 + 0.1 * torch.randn(c, generator=g)
        s[prefix + ".bias"] = 0.1 * torch.randn(c, generator=g)
        s[prefix + ".running_mean"] = 0.1 * torch.randn(c, generator=g)
        s[prefix + ".running_var"] = 0.5 + torch.rand(c, generator=g)

    s["resnet.conv1.weight"] = _uniform(g, (32, 1, 3, 3), 9, 1.7)
    bn("resnet.bn1", 32)
    in_planes = 32
    for i, (planes, blocks, stride) in enumerate(((32, 3, 1), (64, 4, 2), (128, 6, 2), (256, 3, 2)), start=1):
        for b, st in enumerate([stride] + [1] * (blocks - 1)):
            pre = f"resnet.layer{i}.{b}."
            s[pre + "conv1.weight"] = _uniform(g, (planes, in_planes, 3, 3), in_planes * 9, 1.7)
            bn(pre + "bn1", planes)
            s[pre + "conv2.weight"] = _uniform(g, (planes, planes, 3, 3), planes * 9, 1.7)
            bn(pre + "bn2", planes)
            if st != 1 or in_planes != planes:
                s[pre + "shortcut.0.weight"] = _uniform(g, (planes, in_planes, 1, 1), in_planes, 1.7)
                bn(pre + "shortcut.1", planes)
            in_planes = planes
    s["resnet.seg_1.weight"] = _uniform(g, (dimension, 5120), 5120)
    s["resnet.seg_1.bias"] = _uniform(g, (dimension,), 5120)
    return s


def synth_audio(num_samples: int, seed: int = 1234, sample_rate: int = 16000, num_speakers: int = 4) -> np.ndarray:
    """Mono float32 stream in [-1,1]: harmonic 'speakers' (f0 in 90..250 Hz, three formant-like
    resonances) gated by a seeded two-state turn-taking chain with some overlap, plus -40 dB noise."""
    if num_samples > 2_000_000:
        # long streams (bench.py): synthesise 2M samples (125 s) and repeat them -- recurring speakers, 7x cheaper
        base = synth_audio(2_000_000, seed, sample_rate, num_speakers)
        return np.tile(base, num_samples // 2_000_000 + 1)[:num_samples]
    rng = np.random.default_rng(seed)
    t = np.arange(num_samples, dtype=np.float64) / sample_rate
    out = np.zeros(num_samples, dtype=np.float64)
    seg_len = int(0.25 * sample_rate)
    n_seg = num_samples // seg_len + 1
    for _ in range(num_speakers):
        f0 = rng.uniform(90, 250)
        formants = rng.uniform([300, 900, 2200], [800, 2200, 3400])
        vib = 1.0 + 0.02 * np.sin(2 * np.pi * rng.uniform(3, 6) * t + rng.uniform(0, 6.28))
        phase = 2 * np.pi * np.cumsum(f0 * vib) / sample_rate
        voice = np.zeros(num_samples)
        for h in range(1, 30):
            fh = f0 * h
            if fh > 3800:
                break
            amp = sum(np.exp(-0.5 * ((fh - fc) / 180.0) ** 2) for fc in formants) + 0.02
            voice += amp / h ** 0.5 * np.sin(h * phase + rng.uniform(0, 6.28))
        voice /= np.max(np.abs(voice)) + 1e-9
        state, gate = rng.random() < 0.4, np.zeros(n_seg)
        for i in range(n_seg):
            if rng.random() < (0.25 if state else 0.12):
                state = not state
            gate[i] = 1.0 if state else 0.0
        gt = np.repeat(gate, seg_len)[:num_samples]
        k = np.hanning(int(0.05 * sample_rate))
        gt = np.convolve(gt, k / k.sum(), mode="same")
        out += rng.uniform(0.25, 0.5) * gt * voice
    out += 10 ** (-40 / 20) * rng.standard_normal(num_samples)
    return np.clip(out, -1, 1).astype(np.float32)


def windows(stream: np.ndarray, n_chunks: int, chunk: int = 80000, step: int = 8000, first: int = 0) -> np.ndarray:
    """Chunk i = samples [step*i, step*i + chunk), exactly as ``rearrange_audio_stream``
    (reference ``src/diart/operators.py:44-100``) emits them."""
    idx = np.arange(chunk)[None, :] + step * (first + np.arange(n_chunks))[:, None]
    return stream[idx]
