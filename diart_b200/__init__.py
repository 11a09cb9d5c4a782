"""diart_b200: B200-native (sm_100a CUDA) implementation of diart's per-chunk diarization hot path
behind diart's own blocks / model-loader API.  See DESIGN.md and INTEGRATION.md."""
from . import blocks, core, features, mapping, models, synth  # noqa: F401

__version__ = "0.1.0"
