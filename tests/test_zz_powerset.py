"""Powerset segmentation models (pyannote/segmentation-3.0 layout; reference ``PowersetAdapter``,
``src/diart/models.py:29-39``): Linear(128, 7) -> log_softmax -> one_hot(argmax) @ mapping.

pyannote.audio is not vendored by the reference (``setup.cfg:34``) and is absent here, so ``Powerset`` is restated in
``oracle/nets.py`` from its published definition (subsets ordered by size, then as ``itertools.combinations`` yields them);
the CPU tests pin that restatement, the GPU tests compare the CUDA decoding with it.  (File name: runs last.)"""
import ctypes as C

import numpy as np
import pytest
import torch

from diart_b200 import _lib, models, synth
from oracle import nets


def test_powerset_mapping_order():
    m = nets.powerset_mapping(3, 2)
    assert m.tolist() == [[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 0], [1, 0, 1], [0, 1, 1]]
    m4 = nets.powerset_mapping(4, 2)
    assert m4.shape == (11, 4)
    assert m4[5:].tolist() == [[1, 1, 0, 0], [1, 0, 1, 0], [1, 0, 0, 1], [0, 1, 1, 0], [0, 1, 0, 1], [0, 0, 1, 1]]
    assert nets.powerset_mapping(3, 3).shape == (8, 3) and nets.powerset_mapping(2, 1).tolist() == [[0, 0], [1, 0], [0, 1]]


def test_to_multilabel_takes_the_first_maximum():
    m = nets.powerset_mapping(3, 2)
    logp = torch.full((1, 3, 7), -5.0)
    logp[0, 0, 4] = -0.1                      # {0, 1}
    logp[0, 1, 0] = -0.1                      # nobody
    logp[0, 2, 2] = logp[0, 2, 6] = -0.1      # tie between {1} and {1, 2}: argmax returns the first
    assert nets.to_multilabel(logp, m)[0].tolist() == [[1, 1, 0], [0, 0, 0], [0, 1, 0]]


def test_powerset_oracle_net_shapes():
    net = nets.make_powerset_segmentation()
    assert net.classifier.out_features == 7
    x = torch.from_numpy(synth.windows(synth.synth_audio(80000 + 8000, seed=99), 2))
    with torch.no_grad():
        y = net(x[:, None, :])
    assert y.shape == (2, 293, 3) and set(y.unique().tolist()) <= {0.0, 1.0}
    assert (y.sum(-1) <= 2).all()             # at most two speakers per frame


def test_set_powerset_argument_errors_without_gpu():
    lib = _lib.lib()
    assert lib.dg_seg_set_powerset(None, 3, 2) == -1
    assert b"dg_seg_set_powerset" in lib.dg_last_error()


@pytest.mark.gpu
def test_powerset_segmentation_matches_oracle(cuda_device):
    net = nets.make_powerset_segmentation()
    x = torch.from_numpy(synth.windows(synth.synth_audio(80000 + 8000 * 7, seed=77), 8))
    taps = {}
    with torch.no_grad():
        ref = net(x[:, None, :], taps)
    top2 = taps["log_probabilities"].topk(2, dim=-1).values
    sure = (top2[..., 0] - top2[..., 1]) > 1e-2        # frames whose arg-max survives float32-level differences of the logits
    assert sure.float().mean() > 0.9
    seg = models.B200PyanNet(net.state_dict(), powerset=(3, 2)).to(cuda_device)
    assert seg.dims(80000) == (293, 3)
    out = seg(x[:, None, :].to(cuda_device)).cpu()
    assert out.shape == ref.shape and set(out.unique().tolist()) <= {0.0, 1.0}
    assert torch.equal(out[sure], ref[sure])
    print(f"powerset: {int(sure.sum())} of {sure.numel()} frames compared, all equal")


@pytest.mark.gpu
def test_powerset_declaration_must_match_the_classifier(cuda_device):
    net = nets.make_powerset_segmentation()
    with pytest.raises((ValueError, _lib.DiartB200Error)):
        models.B200PyanNet(net.state_dict(), powerset=(4, 2)).to(cuda_device)     # 11 classes declared, 7 outputs
    plain = models.B200PyanNet(net.state_dict()).to(cuda_device)                 # not declared: 7 sigmoid outputs
    assert plain.dims(80000) == (293, 7)
