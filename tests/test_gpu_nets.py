"""Parity of the CUDA networks (through the C ABI) against the oracle restatement (oracle/nets.py).

Tolerances (float32 everywhere, different summation order only): segmentation scores 1e-4 absolute (measured
<= 3e-5; the float32 oracle itself is 0.84e-4 from a float64 evaluation of this synthetic net), unit-norm embeddings 1e-4
absolute (measured <= 2e-5), raw embeddings 3e-4 relative (the synthetic embedding has a ~6x cancellation between its raw
and centred components, see oracle/calibrate.py)."""
import numpy as np
import pytest
import torch

from diart_b200 import models

pytestmark = pytest.mark.gpu

SEG_TOL = 1e-4
EMB_TOL = 1e-4


def _osp(seg, gamma=3, beta=10):
    probs = torch.softmax(beta * seg, dim=-1)
    w = torch.pow(seg, gamma) * torch.pow(probs, gamma)
    w[w < 1e-8] = 1e-8
    return w


@pytest.fixture(scope="module")
def cuda_nets(cuda_device, oracle_nets):
    seg_o, emb_o = oracle_nets
    seg = models.B200PyanNet(seg_o.state_dict()).to(cuda_device)
    emb = models.B200XVectorSincNet(emb_o.state_dict(), "3.1").to(cuda_device)
    return seg, emb


def test_segmentation_matches_oracle(cuda_nets, oracle_nets, audio_batch, cuda_device):
    seg_c, _ = cuda_nets
    seg_o, _ = oracle_nets
    x = audio_batch[:4]
    with torch.no_grad():
        ref = seg_o(x[:, None, :])
    out = seg_c(x[:, None, :].to(cuda_device)).cpu()
    assert out.shape == ref.shape == (4, 293, 3)
    err = (out - ref).abs().max().item()
    print("segmentation max abs err", err)
    assert err < SEG_TOL


def test_segmentation_batch_invariance(cuda_nets, audio_batch, cuda_device):
    """a chunk's scores do not depend on its position in the batch or on the batch size"""
    seg_c, _ = cuda_nets
    x = audio_batch.to(cuda_device)
    full = seg_c(x[:, None, :])
    one = seg_c(x[5:6, None, :])
    three = seg_c(x[3:6, None, :])
    assert torch.equal(full[5], one[0])
    assert torch.equal(full[3:6], three)


def test_embedding_fused_matches_oracle(cuda_nets, oracle_nets, audio_batch, cuda_device):
    _, emb_c = cuda_nets
    seg_o, emb_o = oracle_nets
    x = audio_batch[:4]
    with torch.no_grad():
        w = _osp(seg_o(x[:, None, :]))
        ref = emb_o.forward_dedup(x[:, None, :], w)
        ref_n = ref / ref.norm(dim=-1, keepdim=True)
    raw = emb_c.forward_fused(x.to(cuda_device), w.to(cuda_device)).cpu()
    nrm = emb_c.forward_fused(x.to(cuda_device), w.to(cuda_device), normalize=True).cpu()
    rel = ((raw - ref).norm(dim=-1) / ref.norm(dim=-1)).max().item()
    err = (nrm - ref_n).abs().max().item()
    print("embedding rel err", rel, "normalised max abs err", err)
    assert err < EMB_TOL
    assert torch.allclose(nrm.norm(dim=-1), torch.ones(4, 3), atol=1e-5)


def test_embedding_reference_call_convention(cuda_nets, oracle_nets, audio_batch, cuda_device):
    """(B*K,1,S) repeated waveforms + (B*K,F) weights, as reference blocks/embedding.py:57-65 passes them:
    identical to the fused path (rows are de-duplicated on the device)."""
    _, emb_c = cuda_nets
    seg_o, emb_o = oracle_nets
    x = audio_batch[:3]
    with torch.no_grad():
        w = _osp(seg_o(x[:, None, :]))
    B, F, K = w.shape
    rep = x[:, None, :].repeat(1, K, 1).reshape(B * K, 1, -1)
    w_rows = w.permute(0, 2, 1).reshape(B * K, F)
    out = emb_c(rep.to(cuda_device), w_rows.to(cuda_device)).reshape(B, K, -1)
    fused = emb_c.forward_fused(x.to(cuda_device), w.to(cuda_device))
    # same trunk; the fused path pools inside TDNN5's epilogue (per-tile partial sums), the row path reads the map back
    assert torch.allclose(out, fused, rtol=2e-5, atol=2e-6)
    with torch.no_grad():
        ref = emb_o(rep, w_rows).reshape(B, K, -1)
    rel = ((out.cpu() - ref).norm(dim=-1) / ref.norm(dim=-1)).max().item()
    assert rel < 3e-4   # raw (un-normalised) embeddings: ~6x cancellation in the synthetic net
    # no weights: plain statistics pooling (mean, unbiased std)
    plain = emb_c(x[:, None, :].to(cuda_device), None).cpu()
    with torch.no_grad():
        ref_plain = emb_o(x[:, None, :], None)
    assert ((plain - ref_plain).norm(dim=-1) / ref_plain.norm(dim=-1)).max().item() < 3e-4


def test_embedding_pool_mode_21(cuda_device, oracle_nets, audio_batch):
    """pyannote.audio 2.1 StatsPool: linear weight resize, no epsilon guards"""
    from oracle import nets

    seg_o, _ = oracle_nets
    emb_o = nets.make_embedding(pool_mode="2.1")
    emb_c = models.B200XVectorSincNet(emb_o.state_dict(), "2.1").to(cuda_device)
    x = audio_batch[:2]
    with torch.no_grad():
        w = _osp(seg_o(x[:, None, :]))
        ref = emb_o.forward_dedup(x[:, None, :], w)
    out = emb_c.forward_fused(x.to(cuda_device), w.to(cuda_device)).cpu()
    assert ((out - ref).norm(dim=-1) / ref.norm(dim=-1)).max().item() < 3e-4


def test_bad_shapes_raise(cuda_nets, cuda_device):
    seg_c, emb_c = cuda_nets
    with pytest.raises(AssertionError):
        seg_c(torch.zeros(2, 2, 80000, device=cuda_device))          # not mono
    with pytest.raises(ValueError):
        seg_c(torch.zeros(1, 1, 100, device=cuda_device))            # too short for the 251-tap filter bank
