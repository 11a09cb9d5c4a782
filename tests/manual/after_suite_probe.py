"""Manual (GPU): what does the FIRST process after the GPU test suite compute?  (bench.py's parity check and smoke() reported a
mismatch of 4e-4 .. 2e-2 three times in that slot and never elsewhere.)  Prints the segmentation error against the oracle for
the block-level call (three times in a row), the fused step, and the tcgen05 GEMM self-test with O(1) and with tiny operands."""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
t_start = time.time()
from diart_b200 import _lib, blocks, models, synth  # noqa: E402
from oracle import nets  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "probe"
device = torch.device("cuda", 0)
lib = _lib.lib()


def gemm(shape):
    diff, rms = C.c_float(), C.c_float()
    _lib.check(lib.dg_selftest_gemm_tc(*shape, C.byref(diff), C.byref(rms)))
    return diff.value / max(rms.value, 1e-30)


first_gemm = gemm((4096, 256, 1, 1, 1024, 0))
seg_o, emb_o = nets.make_segmentation(), nets.make_embedding()
stream = synth.synth_audio(80000 + 8000 * 3, seed=1234)
x = torch.from_numpy(synth.windows(stream, 4))
with torch.no_grad():
    ref = seg_o(x[:, None, :])
seg_m = models.SegmentationModel(models.B200SegmentationLoader(seg_o.state_dict()))
blk = blocks.SpeakerSegmentation(seg_m, device)
errs = []
for _ in range(3):
    y = blk.forward_device(x.to(device))
    torch.cuda.synchronize()
    errs.append((y.cpu() - ref).abs().max().item())
config = blocks.SpeakerDiarizationConfig(
    segmentation=models.SegmentationModel(models.B200SegmentationLoader(seg_o.state_dict())),
    embedding=models.EmbeddingModel(models.B200EmbeddingLoader(emb_o.state_dict())), device=device)
pipe = blocks.SpeakerDiarization(config)
perr = []
for _ in range(3):
    seg, emb, maps = pipe.device_step(x.to(device))
    torch.cuda.synchronize()
    perr.append((seg.cpu() - ref).abs().max().item())
    pipe.reset()
import hashlib  # noqa: E402

md5 = lambda a: hashlib.md5(np.ascontiguousarray(a).tobytes()).hexdigest()[:10]
with torch.no_grad():
    ref64 = seg_o.double()(x[:, None, :].double()).float()
print(f"{tag}: md5 of the oracle output {md5(ref.numpy())}, of the float64 oracle {md5(ref64.numpy())}, of the CUDA block output "
      f"{md5(y.cpu().numpy())}, of the fused step {md5(seg.cpu().numpy())}; CUDA vs float64 oracle {(y.cpu() - ref64).abs().max().item():.2e}, "
      f"float32 oracle vs float64 oracle {(ref - ref64).abs().max().item():.2e}, torch threads {torch.get_num_threads()}", flush=True)
print(f"{tag}: started {t_start:.1f}; gemm rel err first {first_gemm:.2e} again {gemm((4096, 256, 1, 1, 1024, 0)):.2e}; "
      f"block seg err x3 {' '.join(f'{e:.2e}' for e in errs)}; fused step seg err x3 {' '.join(f'{e:.2e}' for e in perr)}", flush=True)
