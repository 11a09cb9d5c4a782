"""Shared-identity mode (extension, BASELINE config 5).  CPU: properties of the merge rule on the oracle simulation.
GPU (one device, ranks simulated by separate clustering handles, records concatenated instead of all-gathered):
the CUDA export / merge / relabel kernels against the oracle simulation -- bit-identical tables, identical maps."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle.clustering import OracleClustering
from oracle.shared_identity import export_record, merge
from oracle.synth_cluster import make_stream

M, D, DELTA = 20, 512, 1.0


def simulate(streams, n_steps, batch):
    """G oracle clusterings sharing identity: -> per-step relabelled maps per rank, final table"""
    G = len(streams)
    clus = [OracleClustering(0.6, 0.3, DELTA, "cosine", M) for _ in range(G)]
    base, base_active = np.zeros((M, D)), set()
    all_maps = [[] for _ in range(G)]
    for step in range(n_steps):
        records, step_maps = [], []
        for r, (seg, emb) in enumerate(streams):
            c = clus[r]
            maps = np.stack([c(seg[i], emb[i])[0] for i in range(step * batch, (step + 1) * batch)])
            step_maps.append(maps)
            records.append(export_record(c.centers, c.active_centers, base, base_active))
        base, base_active, relabels = merge(records, base, base_active, DELTA)
        for r in range(G):
            m = step_maps[r].copy()
            m[m >= 0] = relabels[r][m[m >= 0]]
            all_maps[r].append(m)
            clus[r].centers, clus[r].active_centers = base.copy(), set(base_active)
    return [np.concatenate(m) for m in all_maps], base, base_active


def test_identical_streams_share_one_identity_space():
    """two ranks hearing the same stream must agree with each other exactly, and almost always with a single-rank
    run: after a merge the shared centroids are twice the single-rank sums (cosine distance is scale free), but
    within a step a rank's own updates weigh half as much, so a marginal decision may differ"""
    seg, emb = make_stream(96, 5, sigma=1.0)
    maps, table, active = simulate([(seg, emb), (seg, emb)], n_steps=6, batch=16)
    single = OracleClustering(0.6, 0.3, DELTA, "cosine", M)
    ref = np.stack([single(s, e)[0] for s, e in zip(seg, emb)])
    assert np.array_equal(maps[0], maps[1])
    assert (maps[0] == ref).all(axis=1).mean() > 0.95
    assert np.array_equal(maps[0][:16], ref[:16])     # the first step is exactly the single-rank run
    assert len(active) == len(single.active_centers)


def test_distinct_streams_merge_and_allocate():
    streams = [make_stream(64, 20 + r, sigma=1.2) for r in range(3)]
    maps, table, active = simulate(streams, n_steps=4, batch=16)
    assert 1 <= len(active) <= M
    for m in maps:
        used = set(m[m >= 0].tolist())
        assert used <= active           # every emitted global speaker exists in the shared table


@pytest.mark.gpu
def test_cuda_merge_matches_oracle_simulation(cuda_device):
    from diart_b200 import _lib
    from diart_b200.blocks import OnlineSpeakerClustering

    G, n_steps, batch = 3, 4, 16
    streams = [make_stream(n_steps * batch, 40 + r, sigma=1.2) for r in range(G)]
    ref_maps, ref_table, ref_active = simulate(streams, n_steps, batch)
    lib = _lib.lib()
    clus = [OnlineSpeakerClustering(0.6, 0.3, DELTA, "cosine", M, device=cuda_device) for _ in range(G)]
    got = [[] for _ in range(G)]
    for step in range(n_steps):
        step_maps, recs = [], []
        for r, (seg, emb) in enumerate(streams):
            sl = slice(step * batch, (step + 1) * batch)
            m, _ = clus[r].step_batch(torch.from_numpy(seg[sl]), torch.from_numpy(emb[sl]))
            step_maps.append(m)
            n = lib.dg_cluster_record_len(clus[r]._h)
            rec = torch.empty(n, dtype=torch.float64, device=cuda_device)
            _lib.check(lib.dg_cluster_export_delta(clus[r]._h, rec.data_ptr(), None))
            recs.append(rec)
        gathered = torch.cat(recs)                       # what the all-gather delivers on every rank
        for r in range(G):
            _lib.check(lib.dg_cluster_merge(clus[r]._h, gathered.data_ptr(), G, r, step_maps[r].data_ptr(),
                                            step_maps[r].numel(), None))
            got[r].append(step_maps[r].cpu().numpy())
    for r in range(G):
        assert np.array_equal(np.concatenate(got[r]), ref_maps[r]), f"rank {r}"
        assert np.array_equal(clus[r].centers, ref_table), f"rank {r}: shared table is not bit-identical"
        assert clus[r].active_centers == ref_active


@pytest.mark.gpu
def test_pipelined_identity_exchange_equals_the_serial_protocol(oracle_nets, cuda_device):
    """dg_pipeline_identity_export / merge inside the three-deep submit / collect flow (exchange ordered on the clustering
    stream) == one step at a time with dg_cluster_export_delta / dg_cluster_merge: identical maps and bit-identical tables.
    Two simulated ranks on one device; the all-gather is a concatenation."""
    from diart_b200 import _lib, synth
    from test_gpu_pipeline import make_pipeline

    lib = _lib.lib()
    G, nb, B = 2, 5, 12
    audio = [synth.synth_audio(80000 + 8000 * (nb * B - 1), seed=900 + r, num_speakers=3) for r in range(G)]
    batches = [[torch.from_numpy(synth.windows(audio[r], B, first=i * B)).to(cuda_device) for i in range(nb)] for r in range(G)]
    # --- serial protocol
    ser = [make_pipeline(oracle_nets, cuda_device) for _ in range(G)]
    want = [[] for _ in range(G)]
    for i in range(nb):
        recs, maps = [], []
        for r in range(G):
            _, _, m = ser[r].device_step(batches[r][i])
            maps.append(m)
            n = lib.dg_cluster_record_len(ser[r].clustering._h)
            rec = torch.empty(n, dtype=torch.float64, device=cuda_device)
            _lib.check(lib.dg_cluster_export_delta(ser[r].clustering._h, rec.data_ptr(), _lib.stream_ptr(cuda_device)))
            recs.append(rec)
        gathered = torch.cat(recs)
        for r in range(G):
            _lib.check(lib.dg_cluster_merge(ser[r].clustering._h, gathered.data_ptr(), G, r, maps[r].data_ptr(), maps[r].numel(),
                                            _lib.stream_ptr(cuda_device)))
            want[r].append(maps[r].cpu().numpy())
    # --- pipelined: submit, exchange, collect two steps later
    pip = [make_pipeline(oracle_nets, cuda_device) for _ in range(G)]
    hs = [p._ensure_fused(80000)[0] for p in pip]
    n = lib.dg_cluster_record_len(pip[0].clustering._h)
    got = [[] for _ in range(G)]
    st = _lib.stream_ptr(cuda_device)
    for i in range(nb):
        recs = []
        for r in range(G):
            pip[r].submit(batches[r][i])
            rec = torch.empty(n, dtype=torch.float64, device=cuda_device)
            _lib.check(lib.dg_pipeline_identity_export(hs[r], rec.data_ptr(), st))
            recs.append(rec)
        gathered = torch.cat(recs)
        for r in range(G):
            _lib.check(lib.dg_pipeline_identity_merge(hs[r], gathered.data_ptr(), G, r, st))
        if i >= 2:
            for r in range(G):
                got[r].append(pip[r].collect()[2].cpu().numpy())
    for r in range(G):
        while len(got[r]) < nb:
            got[r].append(pip[r].collect()[2].cpu().numpy())
    torch.cuda.synchronize()
    for r in range(G):
        for i in range(nb):
            assert np.array_equal(want[r][i], got[r][i]), f"rank {r} step {i}"
        assert np.array_equal(ser[r].clustering.centers, pip[r].clustering.centers)
    assert np.array_equal(pip[0].clustering.centers, pip[1].clustering.centers)
