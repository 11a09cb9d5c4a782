"""Manual (GPU): ONE serial 256-window pipeline step between cudaProfilerStart / Stop, for
    DG_TRACE_LAUNCHES=1 DG_NO_OVERLAP=1 ncu --profile-from-start off --set full ... python tests/manual/one_step.py 2> trace.log
(tools/ncu_summary.py joins the launch list with the `dg-trace` lines).  Two warm steps run first (allocations, stream form of
the sinc layer verified, clustering table populated)."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from diart_b200 import blocks, models, synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
embedding = sys.argv[2] if len(sys.argv) > 2 else "xvector"
device = torch.device("cuda", 0)
config = blocks.SpeakerDiarizationConfig(
    segmentation=models.SegmentationModel(models.B200SegmentationLoader(synth.segmentation_state())),
    embedding=models.EmbeddingModel(models.B200EmbeddingLoader(
        synth.wespeaker_state() if embedding == "wespeaker" else synth.embedding_state())),
    device=device)
pipe = blocks.SpeakerDiarization(config)
stream = synth.synth_audio(80000 + 8000 * (3 * B - 1), seed=1234)
batches = [torch.from_numpy(synth.windows(stream, B, first=j * B)).to(device) for j in range(3)]
for j in range(2):
    pipe.device_step(batches[j])
torch.cuda.synchronize()
print("dg-trace-begin", file=sys.stderr, flush=True)
torch.cuda.profiler.start()
pipe.device_step(batches[2])
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("dg-trace-end", file=sys.stderr, flush=True)
