"""Manual (GPU): per-phase SM-clock timing of the tcgen05 recurrence, `DG_LSTM_TIMING=1 python tests/manual/lstm_timing.py [batch]`
(the library prints the table on stderr), followed by the segmentation error against the oracle for the same shape."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
os.environ.setdefault("DG_LSTM_TIMING", "1")

from diart_b200 import blocks, models, synth  # noqa: E402
from oracle import nets  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
device = torch.device("cuda", 0)
seg_o = nets.make_segmentation()
seg = blocks.SpeakerSegmentation(models.SegmentationModel(models.B200SegmentationLoader(seg_o.state_dict())), device)
stream = synth.synth_audio(80000 + 8000 * (B - 1), seed=7)
x = torch.from_numpy(synth.windows(stream, B)).to(device)
y = seg.forward_device(x)
torch.cuda.synchronize()
with torch.no_grad():
    ref = seg_o(x[:4, None, :].cpu())
print(f"rows={os.environ.get('DG_LSTM_ROWS', 'default')} cells={os.environ.get('DG_LSTM_CELLS', 'default')} B={B}: "
      f"seg max abs err (4 windows) {(y[:4].cpu() - ref).abs().max().item():.2e}")
