"""Pins the oracle and the host-side mirrors against the reference's OWN code, imported from
/root/reference/src (authoring container only; skipped on the GPU box where the tree is absent)."""
import numpy as np
import pytest
import torch

from oracle import ref_import

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="reference tree not present")


@pytest.fixture(scope="module")
def ref():
    return ref_import.load()


def test_oracle_clustering_equals_reference_class(ref):
    from diart_b200.core import SlidingWindow, SlidingWindowFeature
    from oracle.clustering import OracleClustering
    from oracle.synth_cluster import make_stream

    sw = SlidingWindow(start=0, duration=5 / 293, step=5 / 293)
    for seed, (M, sigma, delta, tau, rho) in enumerate([(20, 1.2, 1.0, 0.6, 0.3), (4, 1.2, 1.0, 0.6, 0.3),
                                                        (6, 3.0, 0.7, 0.6, 0.2), (20, 0.5, 0.3, 0.6, 0.3)]):
        seg, emb = make_stream(250, 100 + seed, sigma=sigma)
        r = ref.clustering.OnlineSpeakerClustering(tau, rho, delta, "cosine", M)
        o = OracleClustering(tau, rho, delta, "cosine", M)
        for i in range(len(seg)):
            out_r = r(SlidingWindowFeature(seg[i], sw), torch.from_numpy(emb[i])).data
            _, out_o = o(seg[i], emb[i])
            assert np.array_equal(out_r, out_o) and np.array_equal(r.centers, o.centers)
            assert r.active_centers == o.active_centers


def test_post_path_blocks_equal_reference(ref):
    """Binarize and DelayedAggregation (host-side 'next' rows) against the reference implementations"""
    import importlib

    from diart_b200 import blocks
    from diart_b200.core import SlidingWindow, SlidingWindowFeature

    ref_agg = importlib.import_module("diart.blocks.aggregation")
    ref_utils = importlib.import_module("diart.blocks.utils")
    rng = np.random.default_rng(3)
    res = 5 / 293
    for latency, n_buf in [(0.5, 1), (2.0, 4), (5.0, 10)]:
        mine = blocks.DelayedAggregation(0.5, latency, "hamming", "loose")
        theirs = ref_agg.DelayedAggregation(0.5, latency, "hamming", "loose")
        assert mine.num_overlapping_windows == theirs.num_overlapping_windows == n_buf
        for first in (0, 7):
            bufs = [SlidingWindowFeature(rng.random((293, 5)), SlidingWindow(start=0.5 * (first + i), duration=res, step=res))
                    for i in range(n_buf)]
            a, b = mine(bufs), theirs(bufs)
            np.testing.assert_allclose(a.data, b.data, rtol=1e-12)
            assert abs(a.sliding_window.start - b.sliding_window.start) < 1e-12
            assert abs(a.sliding_window.step - b.sliding_window.step) < 1e-12
            ann_a, ann_b = blocks.Binarize(0.6)(a), ref_utils.Binarize(0.6)(b)
            assert ann_a.to_rttm() == ann_b.to_rttm()
    for strategy, mode in [("mean", "strict"), ("first", "center")]:
        mine, theirs = blocks.DelayedAggregation(0.5, 1.5, strategy, mode), ref_agg.DelayedAggregation(0.5, 1.5, strategy, mode)
        bufs = [SlidingWindowFeature(rng.random((293, 2)), SlidingWindow(start=0.5 * (3 + i), duration=res, step=res)) for i in range(3)]
        np.testing.assert_allclose(mine(bufs).data, theirs(bufs).data, rtol=1e-12)


def test_formatter_equals_reference(ref):
    from diart_b200.core import SlidingWindow, SlidingWindowFeature
    from diart_b200.features import TemporalFeatureFormatter

    swf = SlidingWindowFeature(np.random.rand(50, 3), SlidingWindow(start=1.5, duration=0.1, step=0.1))
    a, b = TemporalFeatureFormatter(), ref.features.TemporalFeatureFormatter()
    assert torch.equal(a.cast(swf), b.cast(swf))
    ra, rb = a.restore_type(torch.ones(1, 25, 2)), b.restore_type(torch.ones(1, 25, 2))
    assert ra.sliding_window.start == rb.sliding_window.start and ra.sliding_window.step == rb.sliding_window.step


def test_preprocessing_blocks_equal_reference(ref):
    """Resample and AdjustVolume (optional pre-processing next to the path) against the reference implementations"""
    import importlib

    from diart_b200 import blocks
    from diart_b200.core import SlidingWindow, SlidingWindowFeature

    ref_utils = importlib.import_module("diart.blocks.utils")
    rng = np.random.default_rng(5)
    batch = torch.from_numpy(rng.standard_normal((3, 8000, 1)).astype(np.float32) * 0.05)
    loud = batch * 100
    for target in (-20.0, 3.0):
        for x in (batch, loud):
            np.testing.assert_allclose(blocks.AdjustVolume(target)(x).numpy(), ref_utils.AdjustVolume(target)(x).numpy(), rtol=1e-6)
    swf = SlidingWindowFeature(batch[0].numpy(), SlidingWindow(start=1.5, duration=1 / 8000, step=1 / 8000))
    a, b = blocks.Resample(8000, 16000)(swf), ref_utils.Resample(8000, 16000)(swf)
    assert a.data.shape == b.data.shape == (16000, 1)
    np.testing.assert_allclose(a.data, b.data, rtol=1e-6, atol=1e-7)
    assert a.sliding_window.start == b.sliding_window.start and a.sliding_window.step == b.sliding_window.step
    np.testing.assert_allclose(blocks.Resample(16000, 8000)(batch).numpy(), ref_utils.Resample(16000, 8000)(batch).numpy(), rtol=1e-6, atol=1e-7)
