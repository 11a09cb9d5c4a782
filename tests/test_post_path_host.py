"""Host half of the device post-path (diart_b200/blocks/post.py): the integer plan that replaces pyannote.core's crop
arithmetic, the turn decoding and the batched audio aggregation, checked on the CPU against the numpy mirrors of the
reference blocks (which tests/test_oracle_vs_reference.py pins to the reference's own DelayedAggregation / Binarize).
The device kernel (csrc/post.cu) is emulated here statement by statement in numpy from the same plan; the kernel
itself is compared with the same mirrors in tests/test_gpu_post.py."""
import numpy as np
import pytest

from diart_b200.blocks.aggregation import DelayedAggregation
from diart_b200.blocks.post import DevicePostPath, aggregate_audio
from diart_b200.blocks.utils import Binarize
from diart_b200.core import SlidingWindow, SlidingWindowFeature

F, K, M = 293, 3, 20


class HostPlanOnly(DevicePostPath):
    """the planner / decoder without a device handle"""

    def __init__(self, step, latency, tau):
        self.step, self.latency, self.tau = float(step), float(latency), float(tau)
        self.F, self.K, self.M = F, K, M
        self.nw = int(round(latency / step))
        self.labels = [f"speaker{g}" for g in range(M)]
        self._hist_start, self._hist_res = np.zeros(0), np.zeros(0)
        self._turns = np.empty(1 << 16, dtype=np.uint32)
        self._h = None


def emulate_kernel(plan, seg_all, map_all, first_index, tau):
    """csrc/post.cu in numpy: seg_all / map_all hold every chunk seen so far, the batch starts at `first_index`"""
    ham = np.hamming(F)
    header, turns = np.zeros((len(plan), 4), np.int32), []
    for c, pl in enumerate(plan):
        nb, nf, first_nf, first_lo = (int(v) for v in pl[:4])
        nfo = first_nf if first_nf > 0 else nf
        act = np.zeros((nfo, M), bool)
        bufs = [first_index + c - (nb - 1) + j for j in range(nb)]
        perm = []
        for b in bufs:
            p = np.zeros((F, M))
            for k, g in enumerate(map_all[b]):
                if g >= 0:
                    p[:, g] = seg_all[b][:, k]
            perm.append(p)
        for fo in range(nfo):
            fa = fo - (nfo - nf)
            if fa < 0:
                act[fo] = perm[0][np.clip(first_lo + fo, 0, F - 1)] > tau
                continue
            num = den = None
            for j in range(nb):
                idx = np.clip(int(pl[4 + j]) + fa, 0, F - 1)
                p = ham[idx] * perm[j][idx]
                num = p if num is None else num + p
                den = ham[idx] if den is None else den + ham[idx]
            act[fo] = num / den > tau
        header[c] = (len(turns), 0, nfo, 0)
        for g in range(M):
            col = np.concatenate([[False], act[:, g], [False]])
            change = np.flatnonzero(col[1:] != col[:-1])
            for on, off in zip(change[0::2], change[1::2]):
                turns.append((g << 20) | (int(on) << 10) | int(off))
        header[c, 1] = len(turns) - header[c, 0]
    return header, np.array(turns, dtype=np.uint32)


@pytest.mark.parametrize("latency,splits", [(0.5, [1, 5, 12]), (2.0, [3, 1, 14]), (5.0, [7, 11]), (1.5, [18])])
def test_plan_and_decoding_equal_the_reference_blocks(latency, splits):
    rng = np.random.default_rng(int(latency * 10))
    step, duration, sr, tau = 0.5, 5.0, 16000, 0.6
    n = sum(splits)
    seg_all = rng.random((n, F, K)).astype(np.float32)
    # smooth the scores in time so that turns are more than one frame long
    seg_all = np.clip(np.cumsum(seg_all - 0.5, axis=1) * 0.2 + 0.55, 0, 1).astype(np.float32)
    base = rng.permutation(M)[:K]
    map_all = np.stack([base if rng.random() < 0.7 else rng.permutation(M)[:K] for _ in range(n)]).astype(np.int32)
    map_all[rng.random((n, K)) < 0.15] = -1
    waves = [SlidingWindowFeature(rng.standard_normal((80000, 1)).astype(np.float32),
                                  SlidingWindow(start=step * i, duration=1 / sr, step=1 / sr)) for i in range(n)]
    # reference order of operations (diarization.py:205-232) with the numpy mirrors
    agg = DelayedAggregation(step, latency, "hamming", "loose")
    audio_agg = DelayedAggregation(step, latency, "first", "center")
    binarize = Binarize(tau)
    pred_buffer, chunk_buffer, want = [], [], []
    res = waves[0].extent.duration / F
    for i in range(n):
        permuted = np.zeros((F, M))
        for k, g in enumerate(map_all[i]):
            if g >= 0:
                permuted[:, g] = seg_all[i][:, k]
        chunk_buffer.append(waves[i])
        pred_buffer.append(SlidingWindowFeature(permuted, SlidingWindow(start=waves[i].extent.start, duration=res, step=res)))
        want.append((binarize(agg(pred_buffer)), audio_agg(chunk_buffer)))
        if len(chunk_buffer) == agg.num_overlapping_windows:
            chunk_buffer, pred_buffer = chunk_buffer[1:], pred_buffer[1:]
    # this repo's split: plan on the host, scores on the "device", decoding on the host
    post = HostPlanOnly(step, latency, tau)
    got, buf, first = [], [], 0
    for b in splits:
        batch = waves[first:first + b]
        starts = np.array([w.extent.start for w in batch])
        plan, out_start, out_res = post.plan(starts, batch[0].extent.duration / F)
        header, turns = emulate_kernel(plan, seg_all, map_all, first, tau)
        anns = post.annotations(header, turns, len(turns), out_start, out_res)
        audio, buf = aggregate_audio(buf, batch, post.nw, step, latency)
        got += list(zip(anns, audio))
        first += b
    assert len(got) == n
    n_lines = 0
    for i, ((a1, w1), (a2, w2)) in enumerate(zip(want, got)):
        assert a1.to_rttm() == a2.to_rttm(), f"chunk {i}"
        assert [(s.start, s.end, t, l) for s, t, l in a1.itertracks(yield_label=True)] == \
               [(s.start, s.end, t, l) for s, t, l in a2.itertracks(yield_label=True)], f"chunk {i}"
        n_lines += a1.to_rttm().count("\n")
        assert np.array_equal(w1.data, w2.data), f"chunk {i}: aggregated audio"
        assert w1.sliding_window.start == w2.sliding_window.start and w1.sliding_window.step == w2.sliding_window.step
    assert n_lines > n // 2


def test_timestamp_shift_and_reset():
    post = HostPlanOnly(0.5, 0.5, 0.6)
    plan, s0, r0 = post.plan(np.array([0.0, 0.5]), 5.0 / F)
    assert plan[0, 2] > 0 and plan[1, 2] == 0 and plan[0, 3] == -1     # only the chunk starting at t = 0 is prepended
    header = np.array([[0, 1, 293, 0], [1, 1, 30, 0]], np.int32)
    turns = np.array([(2 << 20) | (3 << 10) | 9, (5 << 20) | (0 << 10) | 30], np.uint32)
    a = post.annotations(header, turns, 2, s0, r0)
    b = post.annotations(header, turns, 2, s0, r0, shift=10.0)
    (sa, _, la), = list(a[0].itertracks(yield_label=True))
    (sb, _, lb), = list(b[0].itertracks(yield_label=True))
    assert la == lb == "speaker2" and abs(sb.start - sa.start - 10.0) < 1e-12 and a[0].modality == "speech" and b[0].modality is None
