"""Manual diagnostic (GPU): repeatability of the segmentation network, checked against the oracle."""
import sys, os, torch, numpy as np
sys.path.insert(0, os.getcwd())
from diart_b200 import models, synth
from oracle import nets
torch.set_num_threads(16)
dev = torch.device('cuda', 0)
seg_o = nets.make_segmentation()
x = torch.from_numpy(synth.windows(synth.synth_audio(80000 + 8000 * 7, seed=1234), 8))
with torch.no_grad():
    ref = seg_o(x[:4, None, :])
seg = models.B200PyanNet(seg_o.state_dict()).to(dev)
xd = x[:4, None, :].to(dev)
outs = []
for i in range(12):
    o = seg(xd).cpu()
    outs.append(o)
    print(i, 'err vs oracle %.3e' % (o - ref).abs().max().item(), 'vs run0 %.3e' % (o - outs[0]).abs().max().item())
x8 = x[:, None, :].to(dev)
o8 = [seg(x8).cpu() for _ in range(6)]
print('B=8 repeat diffs', [(o - o8[0]).abs().max().item() for o in o8])
