"""Host-side mirror of the reference interface: type plumbing, result objects, post-path blocks, and the
'no silent CPU fallback' rule."""
import numpy as np
import pytest
import torch

from diart_b200 import _lib, blocks, models
from diart_b200.core import Annotation, Segment, SlidingWindow, SlidingWindowFeature
from diart_b200.features import TemporalFeatureFormatter
from diart_b200.mapping import SpeakerMap


def test_formatter_round_trips_like_the_reference():
    f = TemporalFeatureFormatter()
    x = np.random.rand(10, 2)
    t = f.cast(x)
    assert t.shape == (1, 10, 2) and t.dtype == torch.float32
    assert isinstance(f.restore_type(t), np.ndarray)
    swf = SlidingWindowFeature(np.random.rand(10, 2), SlidingWindow(start=2.0, duration=0.5, step=0.5))
    t = f.cast(swf)
    back = f.restore_type(torch.zeros(1, 20, 3))
    assert isinstance(back, SlidingWindowFeature) and back.sliding_window.start == 2.0
    assert abs(back.sliding_window.step - 0.25) < 1e-12              # 5 s / 20 frames
    assert f.restore_type(f.cast(torch.zeros(4, 10, 2))).shape == (4, 10, 2)
    with pytest.raises(ValueError):
        f.cast([1, 2, 3])
    with pytest.raises(AssertionError):
        f.cast(np.zeros(5))


def test_speaker_map_result_object():
    m = SpeakerMap(np.array([2, -1, 0]), 5)
    assert m.valid_assignments() == ([0, 2], [2, 0]) and m.to_dict() == {0: 2, 2: 0}
    assert m.is_source_speaker_mapped(0) and not m.is_source_speaker_mapped(1)
    scores = np.arange(12, dtype=np.float32).reshape(4, 3)
    out = m.apply(scores)
    assert out.dtype == np.float64 and out.shape == (4, 5)
    assert np.array_equal(out[:, 2], scores[:, 0]) and np.array_equal(out[:, 0], scores[:, 2]) and not out[:, 1].any()
    assert m.mapping_matrix[1].min() == 1e10 and m.mapping_matrix[0, 2] == 0


def test_binarize_turns_at_frame_middles():
    res = 0.1
    data = np.zeros((10, 3))
    data[0:3, 0] = 0.9            # active from the very first frame
    data[4:6, 1] = 0.9
    data[8:10, 1] = 0.9           # still active at the last frame
    swf = SlidingWindowFeature(data, SlidingWindow(start=5.0, duration=res, step=res))
    ann = blocks.Binarize(0.5)(swf)
    got = sorted((round(s.start, 6), round(s.end, 6), lab) for s, _, lab in ann.itertracks(yield_label=True))
    assert got == [(5.05, 5.35, "speaker0"), (5.45, 5.65, "speaker1"), (5.85, 6.05, "speaker1")]
    assert "SPEAKER <NA> 1 5.050 0.300 <NA> <NA> speaker0 <NA> <NA>" in ann.to_rttm()


def test_delayed_aggregation_latency_equals_step():
    agg = blocks.DelayedAggregation(step=0.5, latency=0.5, strategy="hamming", cropping_mode="loose")
    assert agg.num_overlapping_windows == 1
    res = 5 / 293
    first = SlidingWindowFeature(np.random.rand(293, 4), SlidingWindow(start=0, duration=res, step=res))
    out = agg([first])
    assert out.data.shape[1] == 4 and abs(out.extent.start) < 1e-9 and abs(out.extent.end - 5.0) < 1e-6
    later = SlidingWindowFeature(np.random.rand(293, 4), SlidingWindow(start=3.0, duration=res, step=res))
    out = agg([later])
    assert 29 <= out.data.shape[0] <= 31 and abs(out.extent.start - 7.5) < 1e-6


def test_no_cpu_fallback():
    net = models.B200PyanNet({})
    with pytest.raises(_lib.DiartB200Error):
        net.to(torch.device("cpu"))
    with pytest.raises(_lib.DiartB200Error):
        net(torch.zeros(1, 1, 80000))             # not on a CUDA device yet
    if not torch.cuda.is_available():
        with pytest.raises(_lib.DiartB200Error):
            net.to(torch.device("cuda"))
        with pytest.raises(_lib.DiartB200Error):
            blocks.OnlineSpeakerClustering(0.6, 0.3, 1.0).step_batch(torch.zeros(1, 293, 3), torch.zeros(1, 3, 512))


def test_loader_plugin_contract():
    """SegmentationModel / EmbeddingModel keep the reference's lazy-loading interface (models.py:112-139)"""
    calls = []

    class Fake:
        def to(self, device):
            calls.append(("to", str(device)))
            return self

        def __call__(self, *args):
            return torch.zeros(1)

    m = models.SegmentationModel(lambda: calls.append("load") or Fake())
    assert not m.is_in_memory()
    m.eval()
    m.to(torch.device("cpu"))
    assert calls == ["load", ("to", "cpu")] and m.is_in_memory()
    e = models.EmbeddingModel(lambda: (lambda w, x=None: np.zeros((2, 4), dtype=np.float32)))
    assert isinstance(e(torch.zeros(2, 1, 10)), torch.Tensor)      # ndarray results become tensors (models.py:262-264)


def test_stream_form_identity_of_the_sinc_layer():
    """The identity the stream form of the sinc layer rests on (diart_b200/csrc/sinc_tc.cu): for window b of a stream,
        maxpool3(|conv(wav_norm(x_b))|) == maxpool3(|A_b * conv(x)[b*hop/10 + t] + (beta - A_b * mu_b) * sum_k h|),
    A_b = gamma * rstd_b, with ONE convolution of the raw stream shared by all overlapping windows (float64 here)."""
    import torch

    from oracle import nets

    S, hop, B = 4000, 400, 5
    g = torch.Generator().manual_seed(3)
    stream = torch.randn((B - 1) * hop + S, generator=g, dtype=torch.float64) * 0.1 + 0.02
    filt = nets.ParamSincFB().filters().detach().double()                    # (80, 1, 251)
    gamma, beta = 1.25, 0.05
    c_stream = torch.nn.functional.conv1d(stream[None, None], filt, stride=10)[0].T      # (P, 80)
    hsum = filt[:, 0, :].sum(-1)
    for b in range(B):
        x = stream[b * hop:b * hop + S]
        mu, var = x.mean(), x.var(unbiased=False)
        rstd = 1.0 / torch.sqrt(var + 1e-5)
        ref = torch.nn.functional.conv1d(((x - mu) * rstd * gamma + beta)[None, None], filt, stride=10)[0].T.abs()
        T0 = ref.shape[0] // 3
        ref = ref[:3 * T0].reshape(T0, 3, 80).amax(1)
        A = gamma * rstd
        got = (A * c_stream[b * hop // 10:b * hop // 10 + 3 * T0] + (beta - A * mu) * hsum).abs().reshape(T0, 3, 80).amax(1)
        assert torch.allclose(got, ref, rtol=1e-10, atol=1e-12)


def test_rttm_sinks(tmp_path):
    """RTTMWriter / PredictionAccumulator (reference sinks.py:25-88): per-chunk turns are appended as they arrive, and
    same-speaker turns closer than the collar are merged when the stream ends; both sinks end with the same RTTM text"""
    from diart_b200 import sinks
    from diart_b200.core import Annotation, Segment

    chunks = []
    for i in range(6):                                        # speaker0: 0.0-0.48 | 0.50-0.98 | ... (20 ms gaps), speaker1 once
        a = Annotation(modality="speech")
        a[Segment(0.5 * i, 0.5 * i + 0.48), 0] = "speaker0"
        if i == 3:
            a[Segment(1.6, 1.9), 1] = "speaker1"
        chunks.append((a, None))
    path = tmp_path / "out.rttm"
    path.write_text("stale\n")
    writer, acc = sinks.RTTMWriter("file1", path), sinks.PredictionAccumulator("file1")
    assert not path.exists()                                  # an existing file is removed up front
    for c in chunks:
        writer.on_next(c)
        acc.on_next(c)
    assert len(path.read_text().splitlines()) == 7            # un-patched: one line per turn
    writer.on_completed()
    acc.on_completed()
    text = path.read_text()
    assert text == acc.get_prediction().to_rttm()
    assert text.splitlines() == ["SPEAKER file1 1 0.000 2.980 <NA> <NA> speaker0 <NA> <NA>",
                                 "SPEAKER file1 1 1.600 0.300 <NA> <NA> speaker1 <NA> <NA>"]
    with pytest.raises(ValueError):
        acc.on_next("not a prediction")
    loaded = sinks.load_rttm(path)
    assert list(loaded) == ["file1"] and loaded["file1"].to_rttm() == text


def test_pipeline_configs_share_the_reference_latency_rules():
    """duration / step / latency of both pipeline configurations (reference blocks/diarization.py:33-60, blocks/vad.py:27-65):
    latency None or "min" = step, "max" = duration, a number is taken as it is; get_file_padding follows from them"""
    seg, emb = models.SegmentationModel(lambda: None), models.EmbeddingModel(lambda: None)
    for make in (lambda **kw: blocks.SpeakerDiarizationConfig(segmentation=seg, embedding=emb, **kw),
                 lambda **kw: blocks.VoiceActivityDetectionConfig(segmentation=seg, **kw)):
        assert make().latency == 0.5 and make(latency="min").latency == 0.5
        assert make(latency="max").latency == 5 and make(latency=2.5, step=0.25).latency == 2.5
        c = make(duration=4, step=0.5, latency=2.0, sample_rate=8000)
        assert (c.duration, c.step, c.latency, c.sample_rate) == (4, 0.5, 2.0, 8000)
        assert c.get_file_padding(file_duration=1.0) == (1.5, 1.5)          # left = duration - (1.0 + right), right = latency - step
        assert isinstance(c, blocks.PipelineConfig)


def test_oracle_accepts_a_reference_value_only_when_it_reproduces():
    """oracle/nets.py: the first float32 evaluation of a torch CPU module is not reproducible on every host;
    the oracle networks evaluate until two consecutive results agree bit for bit"""
    from oracle import nets

    class Flaky(nets._Reproducible, torch.nn.Module):
        def __init__(self, wrong_calls):
            super().__init__()
            self.calls, self.wrong_calls = 0, wrong_calls

        def forward(self, x):
            self.calls += 1
            return x + (1e-3 if self.calls in self.wrong_calls else 0.0)

    x = torch.ones(3)
    with torch.no_grad():
        m = Flaky({1})
        assert torch.equal(m(x), x) and m.calls == 3                 # first call off: second and third agree
        m = Flaky(set())
        assert torch.equal(m(x), x) and m.calls == 2
        m = Flaky({1, 3, 5, 7})
        with pytest.raises(RuntimeError):
            m(x)
        nets.STABLE = False
        try:
            m = Flaky({1})
            assert not torch.equal(m(x), x) and m.calls == 1           # the timing legs of bench.py: one evaluation
        finally:
            nets.STABLE = True


def test_ncu_summary_joins_launches_with_trace_tags(tmp_path):
    """tools/ncu_summary.py: launch list of `ncu --page raw --csv` + the library's dg-trace lines -> per-tag table and traffic.json"""
    import csv
    import json
    import subprocess
    import sys

    cols = ["ID", "Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors_srcunit_tex.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active"]
    units = ["", "", "", "", "us", "%", "Mbyte", "Mbyte", "sector", "%"]
    rows = [["0", "void at::native::fill(float)", "(128, 1, 1)", "(8, 1, 1)", "1.0", "0", "0", "0", "0", "1"],
            ["1", "gemm_tc2_kernel<256, 0>(TcArgs)", "(192, 1, 1)", "(148, 1, 1)", "100.0", "50", "20", "250", "1000", "9"],
            ["2", "lstm_tc3_kernel<1, 0, 4, 16>()", "(544, 1, 1)", "(32, 1, 1)", "400.0", "30", "300", "60", "2000", "45"],
            ["3", "gemm_tc2_kernel<256, 0>(TcArgs)", "(192, 1, 1)", "(148, 1, 1)", "120.0", "54", "80", "250", "3000", "9"]]
    raw = tmp_path / "raw.csv"
    with open(raw, "w", newline="") as f:
        w = csv.writer(f)
        w.writerows([cols, units] + rows)
    trace = tmp_path / "trace.log"
    trace.write_text("dg-trace warm 10 12\ndg-trace-begin\ndg-trace lstm_inproj 40 41\ndg-trace lstm_rec 41 42\n"
                     "dg-trace lstm_inproj 42 43\ndg-trace-end\n")
    tool = __import__("os").path.join(__import__("os").path.dirname(__file__), "..", "tools", "ncu_summary.py")
    out = subprocess.run([sys.executable, tool, str(raw), str(trace), str(tmp_path / "step")], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    traffic = json.loads((tmp_path / "step.json").read_text())
    assert traffic["lstm_inproj"] == pytest.approx((20 + 250 + 80 + 250) * 1e6 / 2) and traffic["lstm_rec"] == pytest.approx(360e6)
    assert traffic["_step"] == pytest.approx(960e6) and traffic["_launches_per_step"] == {"lstm_inproj": 2, "lstm_rec": 1}
    table = (tmp_path / "step.md").read_text()
    assert "`lstm_rec`" in table and "| 400.0 |" in table and (tmp_path / "step_raw_selected.csv").exists()
