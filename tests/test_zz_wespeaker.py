"""Variant B of the embedding row (SURVEY.md 8(a) A8'): pyannote/wespeaker-voxceleb-resnet34-LM on the GPU -- kaldi fbank as a
tcgen05 GEMM over an overlapping-row view of the waveform, ResNet34 as shifted-window Conv2d GEMMs on zero-padded channels-last
maps (csrc/resnet.cu, TC_CONV2D epilogue of csrc/gemm_tc.cu) -- against the oracle restatement oracle.nets.WeSpeakerResNet34
(pinned by parameter count and map sizes in tests/test_oracle_golden.py).  Bars: stage by stage 2e-4 of the map's scale,
unit-norm embeddings 1e-4 (the bar of variant A)."""
import ctypes as C

import numpy as np
import pytest
import torch

from diart_b200 import _lib, blocks, models, synth
from oracle import nets
from oracle.clustering import OracleClustering

N = 3


@pytest.fixture(scope="module")
def wespeaker():
    torch.set_num_threads(16)
    return nets.make_wespeaker()


@pytest.fixture(scope="module")
def audio():
    return torch.from_numpy(synth.windows(synth.synth_audio(80000 + 8000 * (N - 1), seed=99), N))


def test_variant_is_recognised_without_gpu():
    lib = _lib.lib()
    assert hasattr(lib, "dg_emb_debug_trunk")
    assert lib.dg_emb_debug_trunk(None, None, 1, 80000, 0, None, 0, None) == -1


@pytest.mark.gpu
def test_fbank_and_every_stage_match_the_oracle(wespeaker, audio, cuda_device):
    from torchaudio.compliance import kaldi

    lib = _lib.lib()
    emb = models.B200EmbeddingLoader(wespeaker.state_dict())().to(cuda_device)
    assert emb.dims(80000) == (63, 256)
    x = audio.to(cuda_device)
    dims = (C.c_int * 4)()

    def stage(stop, numel):
        out = np.empty(numel, np.float32)
        _lib.check(lib.dg_emb_debug_trunk(emb.handle, x.data_ptr(), N, 80000, stop, out.ctypes.data, out.size, dims))
        return out.reshape(tuple(dims))

    with torch.no_grad():
        raw = torch.stack([kaldi.fbank(w[None, :] * (1 << 15), num_mel_bins=80, frame_length=25, frame_shift=10, dither=0.0,
                                       sample_frequency=16000, window_type="hamming", use_energy=False) for w in audio])
        got = stage(-2, N * 498 * 80)[..., 0]
        err = np.abs(got - raw.numpy()).max()
        print(f"log-mel max abs err {err:.2e} (values in [{raw.min():.1f}, {raw.max():.1f}])")
        assert err < 2e-3, "fbank"         # float32 FFT vs split-precision DFT on int16-scaled audio: log domain
        r = wespeaker.resnet
        fb = wespeaker.compute_fbank(audio[:, None, :])
        cur = fb.permute(0, 2, 1).unsqueeze(1)
        cur = torch.relu(r.bn1(r.conv1(cur)))
        want = {-1: cur}
        bi = 0
        for layer in (r.layer1, r.layer2, r.layer3, r.layer4):
            for blk in layer:
                cur = blk(cur)
                want[bi] = cur
                bi += 1
        failed = []
        for stop in (-1, 0, 1, 2, 3, 4, 7, 8, 12, 13, 14, 15):
            ref = want[stop].permute(0, 3, 2, 1).numpy()            # (N, C, mel, time) -> (N, time, mel, C)
            got = stage(stop, ref.size)
            assert got.shape == ref.shape, (stop, got.shape, ref.shape)
            scale = np.abs(ref).max()
            err = np.abs(got - ref).max() / scale
            print(f"after {'stem' if stop < 0 else 'block %d' % stop}: shape {ref.shape}, max abs err / scale {err:.2e}")
            if not err < 2e-4:
                failed.append((stop, float(err)))
        assert not failed, f"stages beyond the bar: {failed}"


@pytest.mark.gpu
def test_wespeaker_embeddings_match_oracle(wespeaker, audio, cuda_device):
    emb = models.B200EmbeddingLoader(wespeaker.state_dict())().to(cuda_device)
    g = torch.Generator().manual_seed(3)
    w = torch.rand((N, 293, 3), generator=g) ** 3
    with torch.no_grad():
        ref = wespeaker.forward_dedup(audio[:, None, :], w)
        ref_n = ref / ref.norm(dim=-1, keepdim=True)
    nrm = emb.forward_fused(audio.to(cuda_device), w.to(cuda_device), normalize=True).cpu()
    err = (nrm - ref_n).abs().max().item()
    print(f"WeSpeaker unit-norm embeddings: max abs err {err:.2e}")
    assert err < 1e-4
    # the reference's call convention: rows repeated once per local speaker, weights (N*K, F)
    rep = audio[:, None, :].repeat(1, 3, 1).reshape(N * 3, 1, -1)
    rows = emb(rep.to(cuda_device), w.permute(0, 2, 1).reshape(N * 3, 293).to(cuda_device)).reshape(N, 3, -1).cpu()
    raw = emb.forward_fused(audio.to(cuda_device), w.to(cuda_device)).cpu()
    assert torch.equal(rows, raw)
    plain = emb(audio[:, None, :].to(cuda_device), None).cpu()
    with torch.no_grad():
        ref_plain = wespeaker(audio[:, None, :], None)
    assert ((plain - ref_plain).norm(dim=-1) / ref_plain.norm(dim=-1)).max().item() < 3e-4


@pytest.mark.gpu
def test_pipeline_with_wespeaker_embedding(wespeaker, oracle_nets, cuda_device):
    """the fused step with variant B behind the same EmbeddingModel loader: scores / embeddings against the oracle networks,
    speaker maps identical to the oracle clustering replayed on them"""
    seg_o, _ = oracle_nets
    config = blocks.SpeakerDiarizationConfig(
        segmentation=models.SegmentationModel(models.B200SegmentationLoader(seg_o.state_dict())),
        embedding=models.EmbeddingModel(models.B200EmbeddingLoader(wespeaker.state_dict())), device=cuda_device)
    pipe = blocks.SpeakerDiarization(config)
    n = 12
    stream = synth.synth_audio(80000 + 8000 * (n - 1), seed=4242, num_speakers=4)
    x = torch.from_numpy(synth.windows(stream, n))
    seg, emb, maps = (t.cpu().numpy() for t in pipe.device_step(x.to(cuda_device)))
    assert emb.shape == (n, 3, 256)
    from oracle.pipeline import osp_block

    with torch.no_grad():
        o_seg = seg_o(x[:, None, :])
        o_emb = wespeaker.forward_dedup(x[:, None, :], osp_block(o_seg))
        o_emb = (o_emb / o_emb.norm(dim=-1, keepdim=True)).numpy()
    assert np.abs(seg - o_seg.numpy()).max() < 1e-4
    print(f"pipeline (variant B) emb max abs err {np.abs(emb - o_emb).max():.2e}")
    assert np.abs(emb - o_emb).max() < 1e-4
    replay = OracleClustering(0.6, 0.3, 1.0, "cosine", 20)
    want = np.stack([replay(s, e)[0] for s, e in zip(seg, emb)])
    assert np.array_equal(maps, want)
    assert np.array_equal(pipe.clustering.centers, replay.centers)
