"""Bit-exact parity of the CUDA clustering step (dg_cluster_step) against the oracle restatement of
reference blocks/clustering.py + mapping.py, on structured synthetic streams that reach every branch.
Bar: identical int32 speaker maps, bit-identical float64 centroids, identical permuted scores."""
import numpy as np
import pytest
import torch

from diart_b200.blocks import OnlineSpeakerClustering
from diart_b200.core import SlidingWindow, SlidingWindowFeature
from oracle.clustering import OracleClustering
from oracle.synth_cluster import make_stream

pytestmark = pytest.mark.gpu

CONFIGS = [  # (max_speakers, sigma, delta, tau, rho, K)
    (20, 1.2, 1.0, 0.6, 0.3, 3),
    (4, 1.2, 1.0, 0.6, 0.3, 3),
    (20, 2.5, 0.8, 0.5, 0.3, 3),
    (6, 3.0, 0.7, 0.6, 0.2, 3),
    (3, 1.0, 1.0, 0.6, 0.3, 3),
    (20, 0.5, 0.3, 0.6, 0.3, 3),
    (20, 1.5, 0.9, 0.6, 0.3, 4),
    (5, 2.0, 0.9, 0.55, 0.25, 4),
]


def _oracle_run(seg, emb, tau, rho, delta, M, metric="cosine"):
    o = OracleClustering(tau, rho, delta, metric, M)
    maps, outs = [], []
    for s, e in zip(seg, emb):
        a, out = o(s, e)
        maps.append(a)
        outs.append(out)
    return o, np.stack(maps), np.stack(outs)


@pytest.mark.parametrize("cfg", list(enumerate(CONFIGS)))
def test_cluster_stream_bit_exact(cfg, cuda_device):
    seed, (M, sigma, delta, tau, rho, K) = cfg
    seg, emb = make_stream(768, seed, K=K, sigma=sigma)
    o, ref_maps, ref_out = _oracle_run(seg, emb, tau, rho, delta, M)
    c = OnlineSpeakerClustering(tau, rho, delta, "cosine", M, device=cuda_device)
    got_maps, got_out = [], []
    # uneven batch sizes: state must carry across calls exactly like the reference's per-chunk loop
    edges = [0, 1, 2, 5, 37, 256, 512, 513, 768]
    for lo, hi in zip(edges[:-1], edges[1:]):
        m, p = c.step_batch(torch.from_numpy(seg[lo:hi]), torch.from_numpy(emb[lo:hi]), want_permuted=True)
        got_maps.append(m.cpu().numpy())
        got_out.append(p.cpu().numpy())
    got_maps, got_out = np.concatenate(got_maps), np.concatenate(got_out)
    bad = np.where((got_maps != ref_maps).any(axis=1))[0]
    assert bad.size == 0, f"first differing chunk {bad[:5]}: {got_maps[bad[0]]} vs {ref_maps[bad[0]]}"
    assert np.array_equal(got_out.astype(np.float64), ref_out)
    assert c.active_centers == o.active_centers
    assert np.array_equal(c.centers, o.centers), "centroids are not bit-identical"


@pytest.mark.parametrize("metric,delta", [("euclidean", 1.3), ("sqeuclidean", 1.7), ("cityblock", 22.0), ("chebyshev", 0.16)])
def test_cluster_other_cdist_metrics(metric, delta, cuda_device):
    """the reference hands `metric` to scipy's cdist (mapping.py:175): the non-cosine metrics on unit-norm embeddings, thresholds
    near the median assigned distance so that every branch (assign, create, re-assign on a full table) is reached"""
    seg, emb = make_stream(512, 5, K=3, sigma=1.5)
    emb = (emb / np.linalg.norm(emb, axis=-1, keepdims=True)).astype(np.float32)
    for M in (20, 4):
        o, ref_maps, _ = _oracle_run(seg, emb, 0.6, 0.3, delta, M, metric)
        c = OnlineSpeakerClustering(0.6, 0.3, delta, metric, M, device=cuda_device)
        got = np.concatenate([c.step_batch(torch.from_numpy(seg[lo:hi]), torch.from_numpy(emb[lo:hi]))[0].cpu().numpy()
                              for lo, hi in ((0, 3), (3, 200), (200, 512))])
        bad = np.where((got != ref_maps).any(axis=1))[0]
        assert bad.size == 0, f"{metric}, M={M}: first differing chunk {bad[:5]}"
        assert c.active_centers == o.active_centers
        assert np.array_equal(c.centers, o.centers), "centroids are not bit-identical"
        assert len(o.active_centers) > 1, "the threshold never created a second speaker"


def test_cluster_single_chunk_api(cuda_device):
    """the reference's per-chunk call: SlidingWindowFeature in, permuted SlidingWindowFeature out"""
    seg, emb = make_stream(40, 11)
    o = OracleClustering(0.6, 0.3, 1.0, "cosine", 20)
    c = OnlineSpeakerClustering(0.6, 0.3, 1.0, "cosine", 20, device=cuda_device)
    assert c.centers is None
    sw = SlidingWindow(start=0, duration=5 / 293, step=5 / 293)
    for s, e in zip(seg, emb):
        out = c(SlidingWindowFeature(s, sw), torch.from_numpy(e))
        _, ref = o(s, e)
        assert out.data.dtype == np.float64 and np.array_equal(out.data, ref)
    assert c.num_known_speakers == len(o.active_centers)
    assert c.num_free_centers == 20 - len(o.active_centers)
    assert c.get_next_center_position() == min(set(range(20)) - o.active_centers)


def test_cluster_state_roundtrip_and_reset(cuda_device):
    seg, emb = make_stream(64, 3)
    a = OnlineSpeakerClustering(0.6, 0.3, 1.0, "cosine", 20, device=cuda_device)
    a.step_batch(torch.from_numpy(seg[:32]), torch.from_numpy(emb[:32]))
    b = OnlineSpeakerClustering(0.6, 0.3, 1.0, "cosine", 20, device=cuda_device)
    b._set_state(a.centers, a.active_centers, True)
    ma, _ = a.step_batch(torch.from_numpy(seg[32:]), torch.from_numpy(emb[32:]))
    mb, _ = b.step_batch(torch.from_numpy(seg[32:]), torch.from_numpy(emb[32:]))
    assert torch.equal(ma, mb) and np.array_equal(a.centers, b.centers)
    a.reset()
    assert a.centers is None and a.active_centers == set()


def test_cluster_host_side_mutators(cuda_device):
    """init_centers / add_center / update keep the reference's semantics (clustering.py:73-118)"""
    c = OnlineSpeakerClustering(0.6, 0.3, 1.0, "cosine", 4, device=cuda_device)
    c.init_centers(8)
    assert c.centers.shape == (4, 8) and c.num_known_speakers == 0
    e = np.arange(16, dtype=np.float64).reshape(2, 8)
    assert c.add_center(e[0]) == 0 and c.add_center(e[1]) == 1
    c.update([(0, 1)], e)
    assert np.array_equal(c.centers[1], e[1] + e[0])
    with pytest.raises(AssertionError):
        c.update([(0, 3)], e)
