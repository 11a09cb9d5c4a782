"""Device-side rearrange_audio_stream (dg_stream, diart_b200/operators.py; reference src/diart/operators.py:44-100): windows
formed from the ring buffer in HBM are bit-identical to the host-stacked windows, for ragged pushes and ring wrap-around, and
the pipeline fed from the stream gives exactly what it gives on the stacked windows."""
import numpy as np
import pytest
import torch

from diart_b200 import _lib, synth
from diart_b200.core import SlidingWindow, SlidingWindowFeature
from diart_b200.operators import DeviceAudioStream
from test_gpu_pipeline import make_pipeline

pytestmark = pytest.mark.gpu


def test_windows_from_the_ring_equal_stacked_windows(cuda_device):
    n = 300
    audio = synth.synth_audio(80000 + 8000 * (n - 1), seed=5)
    st = DeviceAudioStream(5, 0.5, 16000, max_windows=32, device=cuda_device)
    rng = np.random.default_rng(0)
    pos, emitted = 0, 0
    assert st.available == 0
    while emitted < n:
        # the reference's sources emit arbitrary block sizes: push until at least one batch is available
        while st.available < min(32, n - emitted) and pos < len(audio):
            k = int(rng.integers(1, 40000))
            block = audio[pos:pos + k]
            st.push(block[None, :] if rng.random() < 0.5 else block)
            pos += len(block)
        b = min(st.available, 32, n - emitted)
        got = st.windows(b).cpu().numpy()
        want = synth.windows(audio, b, first=emitted)
        assert np.array_equal(got, want), f"windows {emitted}..{emitted + b}"
        assert np.array_equal(st.host_window(emitted + b - 1)[:, 0], want[-1])
        emitted += b
    with pytest.raises(ValueError):
        st.windows(1)                                   # nothing left
    with pytest.raises(ValueError):
        st.push(np.zeros((2, 5), np.float32))


def test_ring_refuses_to_overwrite_unconsumed_audio(cuda_device):
    st = DeviceAudioStream(5, 0.5, 16000, max_windows=4, device=cuda_device)
    with pytest.raises(ValueError):
        for _ in range(100):
            st.push(np.zeros(16000, np.float32))


def test_pipeline_on_the_stream_equals_pipeline_on_windows(oracle_nets, cuda_device):
    """call_stream / submit_stream (no window upload, stream-form sinc layer without the overlap check) == __call__ /
    submit on the windows rearrange_audio_stream would have emitted"""
    n, sr = 44, 16000
    audio = synth.synth_audio(80000 + 8000 * (n - 1), seed=4242, num_speakers=4)
    a, b = make_pipeline(oracle_nets, cuda_device, latency=1.5), make_pipeline(oracle_nets, cuda_device, latency=1.5)
    chunks = [SlidingWindowFeature(audio[8000 * i:8000 * i + 80000, None], SlidingWindow(start=0.5 * i, duration=1 / sr, step=1 / sr))
              for i in range(n)]
    want = a(chunks[:20]) + a(chunks[20:])
    st = DeviceAudioStream(5, 0.5, sr, max_windows=24, device=cuda_device)
    st.push(audio[:80000 + 8000 * 19])
    got = b.call_stream(st)
    st.push(audio[80000 + 8000 * 19:])
    got += b.call_stream(st, 24)
    assert len(got) == n
    for i, ((a1, w1), (a2, w2)) in enumerate(zip(want, got)):
        assert a1.to_rttm() == a2.to_rttm(), f"chunk {i}"
        assert np.array_equal(w1.data, w2.data) and w1.sliding_window.start == w2.sliding_window.start, f"chunk {i}"
    assert np.array_equal(a.clustering.centers, b.clustering.centers)
    # pipelined form through the C ABI
    lib = _lib.lib()
    c, d = make_pipeline(oracle_nets, cuda_device), make_pipeline(oracle_nets, cuda_device)
    hc, F, K, D = c._ensure_fused(80000)
    hd = d._ensure_fused(80000)[0]
    st.reset()
    st.push(audio)
    outs = []
    for i in range(2):
        _lib.check(lib.dg_pipeline_submit_stream(hc, st.handle, 22))
    for i in range(2):
        s, e, m = np.empty((22, F, K), np.float32), np.empty((22, K, D), np.float32), np.empty((22, K), np.int32)
        _lib.check(lib.dg_pipeline_collect_host(hc, s.ctypes.data, e.ctypes.data, m.ctypes.data))
        outs.append((s, e, m))
    for i in range(2):
        x = np.ascontiguousarray(synth.windows(audio, 22, first=22 * i))
        s, e, m = np.empty((22, F, K), np.float32), np.empty((22, K, D), np.float32), np.empty((22, K), np.int32)
        _lib.check(lib.dg_pipeline_step_host(hd, x.ctypes.data, 22, 80000, s.ctypes.data, e.ctypes.data, m.ctypes.data, None))
        assert np.array_equal(s, outs[i][0]) and np.array_equal(e, outs[i][1]) and np.array_equal(m, outs[i][2]), f"batch {i}"
