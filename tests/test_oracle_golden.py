"""The oracle (oracle/) against the committed golden vectors, which were produced by the REFERENCE'S
OWN code in the authoring container (oracle/make_golden.py).  Runs without a GPU and without
/root/reference."""
import hashlib
import os

import numpy as np
import pytest
import torch

from diart_b200 import synth
from oracle import nets
from oracle.clustering import OracleClustering
from oracle.pipeline import OraclePipeline, normalize_embeddings, osp_block, overlapped_speech_penalty
from oracle.synth_cluster import make_stream

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_parameter_and_frame_counts(oracle_nets):
    seg, emb = oracle_nets
    assert nets.n_params(seg) == 1_472_749      # pyannote/segmentation
    assert nets.n_params(emb) == 4_346_366      # pyannote/embedding
    x = torch.zeros(1, 1, 80000)
    with torch.no_grad():
        assert seg(x).shape == (1, 293, 3)
        assert emb.trunk(x).shape == (1, 1500, 279)


def test_clustering_traces_match_reference():
    g = np.load(os.path.join(GOLD, "cluster_traces.npz"))
    branches = set()
    for cfg in g["configs"]:
        seed, M, sigma, delta, tau, rho, K, n = cfg
        seed, M, K, n = int(seed), int(M), int(K), int(n)
        seg, emb = make_stream(n, seed, K=K, sigma=sigma)
        o = OracleClustering(tau, rho, delta, "cosine", M)
        maps = np.stack([o(s, e)[0] for s, e in zip(seg, emb)])
        assert np.array_equal(maps, g[f"maps_{seed}"].astype(np.int32)), f"config {seed}"
        assert hashlib.sha256(o.centers.tobytes()).hexdigest() == str(g[f"centers_sha_{seed}"])
        assert sorted(o.active_centers) == list(g[f"active_{seed}"])
        if len(o.active_centers) == M:
            branches.add("full")
        if (maps == -1).any():
            branches.add("unmapped")
    assert branches == {"full", "unmapped"}


def test_functional_known_answers():
    g = np.load(os.path.join(GOLD, "functional_kats.npz"))
    seg, emb = torch.from_numpy(g["seg"]), torch.from_numpy(g["emb"])
    for name, (gamma, beta) in {"osp_3_10": (3, 10), "osp_2_5": (2, 5), "osp_2p5_7": (2.5, 7)}.items():
        np.testing.assert_allclose(overlapped_speech_penalty(seg.clone(), gamma, beta).numpy(), g[name], rtol=1e-6)
    np.testing.assert_allclose(osp_block(seg.clone(), 3, 10, True).numpy(), g["osp_norm"], rtol=1e-6)
    assert (g["osp_norm"][1, :, 1] == np.float32(1e-8)).all()       # min == max column -> nan_to_num(1e-8)
    assert g["osp_3_10"].min() == np.float32(1e-8)                   # clamp branch
    np.testing.assert_allclose(normalize_embeddings(emb, 1).numpy(), g["normalize_1"], rtol=1e-6)
    np.testing.assert_allclose(normalize_embeddings(emb, 2.5).numpy(), g["normalize_2p5"], rtol=1e-6)


def test_pipeline_nets_match_reference_blocks(oracle_nets):
    """oracle/pipeline.py (restating diarization.py:177-188, embedding.py:51-68) == the reference's own
    SpeakerSegmentation + OverlapAwareSpeakerEmbedding blocks run over the same networks"""
    g = np.load(os.path.join(GOLD, "nets.npz"))
    seg_net, emb_net = oracle_nets
    stream = synth.synth_audio(80000 + 8000 * 7, seed=1234)
    x = torch.from_numpy(synth.windows(stream, 8)[:2])
    pipe = OraclePipeline(seg_net, emb_net, as_reference=True)
    seg, emb = pipe.nets(x)
    # same code, possibly another CPU / thread count: float32 re-association noise only
    assert np.abs(seg.numpy() - g["seg"]).max() < 2e-5
    assert np.abs(emb.numpy() - g["emb"]).max() < 1e-4
    dedup = OraclePipeline(seg_net, emb_net, as_reference=False).nets(x)[1]
    assert np.abs(dedup.numpy() - g["emb"]).max() < 1e-4   # trunk-once == K-fold repeat


def test_numpy_reduction_semantics():
    """the clustering kernel mirrors np.max / np.mean over axis 0 of a float32 (F,K) array:
    the mean is a float32 running sum in frame order followed by one float32 division"""
    rng = np.random.default_rng(0)
    seg = rng.random((293, 3)).astype(np.float32)
    run = np.zeros(3, dtype=np.float32)
    for f in range(293):
        run = (run + seg[f]).astype(np.float32)
    assert np.array_equal(np.mean(seg, axis=0), run / np.float32(293))
    assert (np.float32(0.6) >= 0.6) and not (np.float32(0.59999996) >= 0.6)   # float32 comparison (weak scalar)


def test_variant_b_oracle_wespeaker_resnet34():
    """Variant B of the embedding row (SURVEY.md 8(a) A8', Appendix A.6) exists as an ORACLE only so far: pinned by the
    published size of WeSpeaker ResNet34 (6.63 M parameters), the feature-map sizes 80x498 -> 10x63, the 5120-d TSTP
    statistics, and the pooling semantics (constant weights = no weights)."""
    import torch

    from diart_b200 import synth
    from oracle import nets

    net = nets.make_wespeaker()
    assert nets.n_params(net) == 6_634_336
    x = torch.from_numpy(synth.windows(synth.synth_audio(80000 + 8000, seed=5), 2))[:, None, :]
    with torch.no_grad():
        fb = net.compute_fbank(x)
        assert fb.shape == (2, 498, 80) and fb.mean(dim=1).abs().max() < 1e-4        # CMN
        maps = net.resnet.maps(fb)
        assert maps.shape == (2, 256, 10, 63)
        assert net.resnet.seg_1.in_features == 5120
        e = net(x)
        e_const = net(x, torch.full((2, 293), 0.37))
        w = torch.rand(2, 293, generator=torch.Generator().manual_seed(1))
        e_w = net(x, w)
    assert e.shape == (2, 256)
    with torch.no_grad():
        w2 = torch.stack([w, 1.0 - w], dim=-1)
        dedup = net.forward_dedup(x, w2)
    assert dedup.shape == (2, 2, 256) and torch.allclose(dedup[:, 0], e_w, rtol=1e-4, atol=1e-5)
    assert torch.allclose(e, e_const, rtol=1e-3, atol=1e-4)       # scale-free weights (+1e-8 terms of the 3.1 StatsPool)
    assert not torch.allclose(e, e_w, rtol=1e-3, atol=1e-4)


def test_fbank_as_two_matrix_products_equals_kaldi():
    """the B200 mapping of variant B's front end (oracle/fbank_linear.py): one [514, 400] operator on an overlapping-row
    view of the waveform, power, one [80, 257] mel matrix, log -- against torchaudio's kaldi.fbank"""
    import torch
    from torchaudio.compliance import kaldi

    from diart_b200 import synth
    from oracle import fbank_linear

    x = synth.synth_audio(80000, seed=11) * 32768.0
    ref = kaldi.fbank(torch.from_numpy(x)[None], num_mel_bins=80, frame_length=25, frame_shift=10, dither=0.0,
                      sample_frequency=16000, window_type="hamming", use_energy=False).numpy()
    got = fbank_linear.fbank(x)
    assert got.shape == ref.shape == (498, 80)
    assert np.abs(got - ref).max() < 2e-3 and np.abs(got - ref).mean() < 1e-4     # log-mel values are O(10)


def test_resnet34_in_shifted_window_gemm_form_equals_the_oracle():
    """the B200 mapping of variant B's trunk (oracle/resnet_gemm_form.py): time-major channels-last maps, 3x3 / 1x1 convolutions
    as 9- / 1-tap shifted-window GEMMs with zero out-of-bounds rows, folded BatchNorm, permuted Linear columns"""
    import torch

    from diart_b200 import synth
    from oracle import nets, resnet_gemm_form

    net = nets.make_wespeaker()
    x = torch.from_numpy(synth.windows(synth.synth_audio(80000 + 8000, seed=5), 2))[:, None, :]
    w = torch.rand(2, 293, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        fb = net.compute_fbank(x)
        for weights in (None, w):
            ref = net.resnet(fb, weights)
            got = resnet_gemm_form.forward(net, fb, weights)
            assert got.shape == ref.shape == (2, 256)
            assert (got - ref).abs().max() < 2e-4 * ref.abs().max()


def test_per_layer_fingerprints_of_the_oracle_networks():
    """golden kind 1 of SURVEY.md 8(c): per-layer tensors of the (restated, third-party) networks on seeded synthetic weights
    and audio -- stored as fingerprints (mean, std, |max|, 32 samples per layer) in tests/golden/net_layers.npz by
    `python oracle/make_golden.py --layers`; any change of the restatement or of the synthetic weights shows up here"""
    import importlib.util

    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(os.path.dirname(GOLD), "..", "oracle", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    g = np.load(os.path.join(GOLD, "net_layers.npz"))
    now = mg.layer_fingerprints()
    assert sorted(k.replace("/", "__") for k in now) == sorted(g.files) and len(g.files) == 19
    for name, fp in now.items():
        ref = g[name.replace("/", "__")]
        scale = max(ref[1], 1e-6)                       # the layer's standard deviation
        assert np.abs(fp - ref).max() < 2e-4 * max(scale, ref[2]), name
