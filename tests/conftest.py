import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def cuda_device():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)


@pytest.fixture(scope="session")
def oracle_nets():
    import torch

    from oracle import nets

    torch.set_num_threads(min(16, os.cpu_count() or 1))
    return nets.make_segmentation(), nets.make_embedding()


@pytest.fixture(scope="session")
def audio_batch():
    """8 consecutive 5 s windows (0.5 s step) of the seeded synthetic stream."""
    import torch

    from diart_b200 import synth

    stream = synth.synth_audio(80000 + 8000 * 7, seed=1234)
    return torch.from_numpy(synth.windows(stream, 8))
