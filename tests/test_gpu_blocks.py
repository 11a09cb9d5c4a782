"""Element-wise blocks and block-level API on the GPU against the golden known answers (produced by the
reference's own functional.py / blocks/embedding.py, see oracle/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from diart_b200 import blocks, models
from diart_b200.core import SlidingWindow, SlidingWindowFeature

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_osp_known_answers(cuda_device):
    g = np.load(os.path.join(GOLD, "functional_kats.npz"))
    seg = torch.from_numpy(g["seg"])
    for name, (gamma, beta) in {"osp_3_10": (3, 10), "osp_2_5": (2, 5), "osp_2p5_7": (2.5, 7)}.items():
        out = blocks.OverlappedSpeechPenalty(gamma, beta, device=cuda_device)(seg)
        assert out.device.type == "cpu"                                   # result comes back where the input lived
        np.testing.assert_allclose(out.numpy(), g[name], rtol=2e-5, atol=1e-12)
    out = blocks.OverlappedSpeechPenalty(3, 10, normalize=True, device=cuda_device)(seg.to(cuda_device))
    assert out.device.type == "cuda"
    np.testing.assert_allclose(out.cpu().numpy(), g["osp_norm"], rtol=5e-5, atol=1e-9)
    assert (out.cpu().numpy()[1, :, 1] == np.float32(1e-8)).all()         # NaN column -> 1e-8
    # numpy / SlidingWindowFeature inputs come back in kind (features.py semantics)
    assert isinstance(blocks.OverlappedSpeechPenalty(device=cuda_device)(g["seg"]), np.ndarray)
    swf = SlidingWindowFeature(g["seg"][0], SlidingWindow(start=0, duration=0.1, step=0.1))
    assert isinstance(blocks.OverlappedSpeechPenalty(device=cuda_device)(swf), SlidingWindowFeature)


def test_normalization_known_answers(cuda_device):
    g = np.load(os.path.join(GOLD, "functional_kats.npz"))
    emb = torch.from_numpy(g["emb"])
    np.testing.assert_allclose(blocks.EmbeddingNormalization(1, cuda_device)(emb).numpy(), g["normalize_1"], rtol=1e-6)
    np.testing.assert_allclose(blocks.EmbeddingNormalization(2.5, cuda_device)(emb).numpy(), g["normalize_2p5"], rtol=1e-6)
    out2d = blocks.EmbeddingNormalization(1, cuda_device)(emb[0])        # (speakers, dim) gains a batch dim
    assert out2d.shape == (1, 3, 16)
    per_spk = torch.tensor([[1.0], [2.0], [3.0]])
    out = blocks.EmbeddingNormalization(per_spk.unsqueeze(0).repeat(3, 1, 1), cuda_device)(emb)
    np.testing.assert_allclose(out.norm(dim=-1).numpy(), np.tile([1.0, 2.0, 3.0], (3, 1)), rtol=1e-5)


def test_blocks_match_golden_networks(cuda_device, oracle_nets, audio_batch):
    """reference-style block calls (batch, samples, channels) -> CPU tensors, against the vectors the
    reference's own blocks produced over the oracle networks"""
    g = np.load(os.path.join(GOLD, "nets.npz"))
    seg_o, emb_o = oracle_nets
    seg_block = blocks.SpeakerSegmentation(models.SegmentationModel(models.B200SegmentationLoader(seg_o.state_dict())),
                                           cuda_device)
    emb_block = blocks.OverlapAwareSpeakerEmbedding(
        models.EmbeddingModel(models.B200EmbeddingLoader(emb_o.state_dict())), 3, 10, 1, False, cuda_device)
    batch = audio_batch[:2, :, None]
    seg = seg_block(batch)
    emb = emb_block(batch, seg)
    assert seg.device.type == "cpu" and emb.device.type == "cpu"
    assert np.abs(seg.numpy() - g["seg"]).max() < 1e-4
    assert np.abs(emb.numpy() - g["emb"]).max() < 1e-4
    # a single SlidingWindowFeature chunk, as StreamingInference feeds it with batch size 1
    swf = SlidingWindowFeature(audio_batch[0].numpy()[:, None], SlidingWindow(start=0, duration=1 / 16000, step=1 / 16000))
    one = seg_block(swf)
    assert isinstance(one, SlidingWindowFeature) and one.data.shape == (293, 3)
    assert np.abs(one.data - g["seg"][0]).max() < 1e-4
    assert abs(one.sliding_window.step - 5 / 293) < 1e-12
