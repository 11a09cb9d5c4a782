"""Manual (GPU): repeatability of the pipelined path under the sequence of legs bench.py runs (pipelined device steps, host-buffer
steps, SpeakerDiarization.__call__ with its sub-batches, reset, three pipelined batches) -- every round must reproduce round 0
bit for bit.  `python tests/manual/stress_parity.py [rounds] [batch]`"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from diart_b200 import _lib, blocks, models, synth  # noqa: E402
from diart_b200.core import SlidingWindow, SlidingWindowFeature  # noqa: E402

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 30
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
CHUNK, STEP, NB = 80000, 8000, 3
device = torch.device("cuda", 0)
lib = _lib.lib()
config = blocks.SpeakerDiarizationConfig(
    segmentation=models.SegmentationModel(models.B200SegmentationLoader(synth.segmentation_state())),
    embedding=models.EmbeddingModel(models.B200EmbeddingLoader(synth.embedding_state())), device=device)
pipe = blocks.SpeakerDiarization(config)
audio = synth.synth_audio(CHUNK + STEP * (NB * B - 1), seed=1234)
host = np.stack([synth.windows(audio, B, first=j * B) for j in range(NB)])
dev = [torch.from_numpy(host[j]).to(device) for j in range(NB)]
pinned = [torch.from_numpy(host[j]).pin_memory() for j in range(NB)]
fused, F, K, D = pipe._ensure_fused(CHUNK)
stream = _lib.stream_ptr(device)
seg_h = [torch.empty((B, F, K)).pin_memory() for _ in range(3)]
emb_h = [torch.empty((B, K, D)).pin_memory() for _ in range(3)]
map_h = [torch.empty((B, K), dtype=torch.int32).pin_memory() for _ in range(3)]
rows = [[np.ascontiguousarray(host[j][b][:, None]) for b in range(B)] for j in range(NB)]
sw_of = lambda n: SlidingWindow(start=0.5 * n, duration=1 / 16000, step=1 / 16000)


def device_steps(n):
    fused = pipe._ensure_fused(CHUNK)[0]       # (pipe.reset() drops the handle: never keep it across a reset)
    for i in range(n):
        _lib.check(lib.dg_pipeline_submit(fused, dev[i % NB].data_ptr(), B, CHUNK, stream))
        if i > 1:
            _lib.check(lib.dg_pipeline_collect(fused, None, None, None, stream))
    for _ in range(min(n, 2)):
        _lib.check(lib.dg_pipeline_collect(fused, None, None, None, stream))


def host_steps(n):
    fused = pipe._ensure_fused(CHUNK)[0]
    for i in range(n):
        _lib.check(lib.dg_pipeline_submit_host(fused, pinned[i % NB].data_ptr(), B, CHUNK))
        if i >= 2:
            j = (i - 2) % 3
            _lib.check(lib.dg_pipeline_collect_host(fused, seg_h[j].data_ptr(), emb_h[j].data_ptr(), map_h[j].data_ptr()))
    for i in range(max(0, n - 2), n):
        j = i % 3
        _lib.check(lib.dg_pipeline_collect_host(fused, seg_h[j].data_ptr(), emb_h[j].data_ptr(), map_h[j].data_ptr()))


def calls(n):
    pipe.reset()
    for i in range(n):
        chunks = [SlidingWindowFeature(rows[i % NB][b], sw_of(i * B + b)) for b in range(B)]
        pipe(chunks)


def fresh_three():
    pipe.reset()
    got = []
    for i in range(NB):
        pipe.submit(dev[i])
        if i > 0:
            got.append(pipe.collect())
    got.append(pipe.collect())
    torch.cuda.synchronize(device)
    return [tuple(t.cpu().numpy() for t in g) for g in got]


first, bad = None, 0
for r in range(rounds):
    device_steps(5 + r % 4)
    host_steps(3 + r % 3)
    if os.environ.get("STRESS_NO_CALLS") != "1":
        calls(1 + r % 2)
    got = fresh_three()
    if first is None:
        first = got
        continue
    for j in range(NB):
        for name, a, b in zip(("seg", "emb", "map"), first[j], got[j]):
            if not np.array_equal(a, b):
                bad += 1
                d = np.abs(a.astype(np.float64) - b.astype(np.float64))
                rows_bad = np.unique(np.where(d.reshape(d.shape[0], -1) > 0)[0])
                print(f"round {r}: batch {j} {name} differs: max {d.max():.3e}, {len(rows_bad)} windows, first {rows_bad[:8]}", flush=True)
print(f"stress_parity: {rounds} rounds at batch {B}: {bad} mismatching tensors")
sys.exit(1 if bad else 0)
