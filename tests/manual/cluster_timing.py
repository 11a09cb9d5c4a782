"""Manual (GPU): per-phase SM-clock timing of the sequential clustering kernel,
`DG_CLUSTER_TIMING=1 python tests/manual/cluster_timing.py` (the library prints the table on stderr)."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
os.environ.setdefault("DG_CLUSTER_TIMING", "1")
from diart_b200.blocks import OnlineSpeakerClustering  # noqa: E402
from oracle.synth_cluster import make_stream  # noqa: E402

device = torch.device("cuda", 0)
seg, emb = make_stream(1024, 3, K=3, sigma=1.2)
c = OnlineSpeakerClustering(0.6, 0.3, 1.0, "cosine", 20, device=device)
for lo in range(0, 1024, 256):
    c.step_batch(torch.from_numpy(seg[lo:lo + 256]).to(device), torch.from_numpy(emb[lo:lo + 256]).to(device))
torch.cuda.synchronize()
print("active speakers", len(c.active_centers))
