"""Stream form of the sinc layer vs the per-window form (diart_b200/csrc/sinc_tc.cu).  Written after the last GPU minutes of round
1 were spent, so this file has not run on hardware yet; both forms did run there against the oracle through
test_gpu_pipeline.py / test_gpu_shapes.py (profiles/r1_stream_form_tests.log).  (File name: runs last.)"""
import numpy as np
import pytest
import torch

from diart_b200 import synth
from test_gpu_pipeline import BATCH, N_CHUNKS, make_pipeline, window_form  # noqa: F401

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def stream():
    return synth.synth_audio(80000 + 8000 * (N_CHUNKS - 1), seed=4242, num_speakers=4)


def test_hop_hint_is_only_a_hint(oracle_nets, stream, cuda_device):
    """a batch that is NOT a run of overlapping windows (here: the windows in reverse order) must give exactly the
    per-window results although the hop hint is set -- the overlap is verified on the device for every batch -- and a
    batch that is one gives the same speaker maps and scores within the parity bar in both forms"""
    a, b = make_pipeline(oracle_nets, cuda_device), window_form(make_pipeline(oracle_nets, cuda_device))
    fwd = synth.windows(stream, BATCH)
    rev = torch.from_numpy(np.ascontiguousarray(fwd[::-1])).to(cuda_device)
    (s1, e1, m1), (s2, e2, m2) = a.device_step(rev), b.device_step(rev)
    assert torch.equal(s1, s2) and torch.equal(e1, e2) and torch.equal(m1, m2)
    a, b = make_pipeline(oracle_nets, cuda_device), window_form(make_pipeline(oracle_nets, cuda_device))   # fresh clustering state
    x = torch.from_numpy(fwd).to(cuda_device)
    (s1, e1, m1), (s2, e2, m2) = a.device_step(x), b.device_step(x)
    assert (s1 - s2).abs().max().item() < 1e-4 and (e1 - e2).abs().max().item() < 1e-4
    assert torch.equal(m1, m2)
