"""WeSpeaker ResNet34 in the form the B200 path will compute it (TEST INFRASTRUCTURE: a CPU prototype of the mapping, checked
against the oracle module ``oracle.nets.WeSpeakerResNet34`` in ``tests/test_oracle_golden.py``).

* maps are ``[item][time w][frequency h][channel]`` (time-major, channels last), so that a map position is one GEMM row
  and the 2560 TSTP features of a time step are contiguous;
* ``Conv2d(k=3, pad=1, stride=s)`` is a shifted-window GEMM with 9 taps: output row ``(w', h')`` reads the rows
  ``(s w' + dw - 1, s h' + dh - 1)``; out-of-range rows contribute zeros (on the device: TMA out-of-bounds fill, traversal
  stride ``s``), K = 9 x Cin ordered tap-major;
* ``BatchNorm2d`` (eval) is folded into a per-channel ``scale, shift`` applied in the GEMM epilogue, then the residual and
  the ReLU;
* the ``Linear(5120, 256)`` columns are permuted once from pyannote's ``(channel, frequency)`` feature order to ``(frequency,
  channel)``.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F


def fold_bn(bn: torch.nn.BatchNorm2d):
    scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    return scale, bn.bias - bn.running_mean * scale


def pack_conv(weight: torch.Tensor) -> torch.Tensor:
    """Conv2d weight [co][ci][kh (frequency)][kw (time)] -> GEMM weight [co][tap = dw * kh + dh][ci] flattened to [co, K]"""
    co, ci, kh, kw = weight.shape
    return weight.permute(0, 3, 2, 1).reshape(co, kw * kh * ci)


def conv_nwhc(x: torch.Tensor, w_packed: torch.Tensor, ksize: int, stride: int) -> torch.Tensor:
    """x [n][w][h][ci] -> [n][w'][h'][co] as ksize*ksize shifted-window GEMM taps (pad = ksize // 2)"""
    n, W, H, ci = x.shape
    pad = ksize // 2
    Wo, Ho = (W + 2 * pad - ksize) // stride + 1, (H + 2 * pad - ksize) // stride + 1
    xp = F.pad(x, (0, 0, pad, pad, pad, pad))                       # zero rows = what TMA out-of-bounds fill supplies
    co = w_packed.shape[0]
    wt = w_packed.reshape(co, ksize * ksize, ci)
    out = x.new_zeros((n, Wo, Ho, co))
    for dw in range(ksize):
        for dh in range(ksize):
            rows = xp[:, dw:dw + stride * (Wo - 1) + 1:stride, dh:dh + stride * (Ho - 1) + 1:stride, :]
            out += rows @ wt[:, dw * ksize + dh, :].T
    return out


def block(x: torch.Tensor, blk) -> torch.Tensor:
    stride = blk.conv1.stride[0]
    s1, b1 = fold_bn(blk.bn1)
    s2, b2 = fold_bn(blk.bn2)
    y = torch.relu(conv_nwhc(x, pack_conv(blk.conv1.weight), 3, stride) * s1 + b1)
    y = conv_nwhc(y, pack_conv(blk.conv2.weight), 3, 1) * s2 + b2
    if len(blk.shortcut):
        ss, bs = fold_bn(blk.shortcut[1])
        x = conv_nwhc(x, pack_conv(blk.shortcut[0].weight), 1, stride) * ss + bs
    return torch.relu(y + x)


def forward(net, fbank: torch.Tensor, weights: Optional[torch.Tensor] = None) -> torch.Tensor:
    """net: oracle.nets.WeSpeakerResNet34; fbank (N, T, 80) mean-normalised -> (N, 256)"""
    r = net.resnet
    x = fbank[:, :, :, None]                                       # [n][w = time][h = mel][1]
    s, b = fold_bn(r.bn1)
    x = torch.relu(conv_nwhc(x, pack_conv(r.conv1.weight), 3, 1) * s + b)
    for layer in (r.layer1, r.layer2, r.layer3, r.layer4):
        for blk in layer:
            x = block(x, blk)
    n, W, H, C = x.shape                                           # (N, 63, 10, 256)
    feats = x.reshape(n, W, H * C).transpose(1, 2)                 # (N, 2560 ordered (h, c), T') as StatsPool expects (N, C, T)
    stats = r.pool(feats, weights)                                 # [mean (h, c) | std (h, c)]
    perm = (torch.arange(C)[None, :] * H + torch.arange(H)[:, None]).reshape(-1)     # position (h, c) <- pyannote column c*H + h
    wlin = torch.cat([r.seg_1.weight[:, :H * C][:, perm], r.seg_1.weight[:, H * C:][:, perm]], dim=1)
    return stats @ wlin.T + r.seg_1.bias
