"""Manual (CPU only, run on the GPU box): is the torch CPU float32 oracle reproducible across processes on this host?
Prints the md5 of three float32 evaluations of the segmentation oracle and of one embedding evaluation, and the distance of the
first one from the float64 evaluation.  `ORACLE_MKLDNN=0` disables torch's oneDNN kernels for the float32 evaluation."""
import hashlib
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
if os.environ.get("ORACLE_MKLDNN") == "0":
    torch.backends.mkldnn.enabled = False
if os.environ.get("ORACLE_THREADS"):
    torch.set_num_threads(int(os.environ["ORACLE_THREADS"]))
from diart_b200 import synth  # noqa: E402
from oracle import nets  # noqa: E402

md5 = lambda a: hashlib.md5(np.ascontiguousarray(a).tobytes()).hexdigest()[:8]
seg_o = nets.make_segmentation()
x = torch.from_numpy(synth.windows(synth.synth_audio(80000 + 8000 * 3, seed=1234), 4))
with torch.no_grad():
    ys = [seg_o(x[:, None, :]) for _ in range(3)]
    y64 = nets.make_segmentation().double()(x[:, None, :].double()).float()
print(sys.argv[1] if len(sys.argv) > 1 else "", "seg float32 x3", " ".join(md5(y.numpy()) for y in ys),
      f"| vs float64 {(ys[0] - y64).abs().max().item():.2e} | threads {torch.get_num_threads()} mkldnn {torch.backends.mkldnn.enabled}", flush=True)
