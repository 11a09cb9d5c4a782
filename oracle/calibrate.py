"""ORACLE helper: calibrates the synthetic weights of diart_b200/synth.py and writes
diart_b200/synth_calib.npz (committed; ~35 KB).

Untrained networks are nearly input-independent, so (1) the classifier rows are rescaled / shifted so
the per-frame logits have unit-ish spread and a mix of active / inactive / short / long local
speakers, (2) the TDNN BatchNorm running statistics are set to the statistics actually observed on a
calibration stream (what training would have done) and (3) the embedding bias removes the common
component so cosine distances spread over (0, 2).  Run once: ``python oracle/calibrate.py``.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diart_b200 import synth  # noqa: E402
from oracle import nets  # noqa: E402

SEG_SEED, EMB_SEED = 4321, 8765


def main():
    torch.set_num_threads(os.cpu_count())
    audio = synth.synth_audio(80000 + 8000 * 127, seed=99, num_speakers=4)
    x = torch.from_numpy(synth.windows(audio, 128))[:, None, :]
    out = {}
    with torch.no_grad():
        seg = nets.make_segmentation(SEG_SEED, calibrated=False)
        logits = torch.logit(seg(x).double().clamp(1e-9, 1 - 1e-9))
        m, sd = logits.mean(dim=(0, 1)), logits.std(dim=(0, 1))
        # per-speaker targets: speaker 0 mostly active & long, 1 sporadic, 2 in between
        tgt_std = torch.tensor([1.8, 1.4, 1.6], dtype=torch.float64)
        tgt_mean = torch.tensor([-0.3, -3.4, -1.9], dtype=torch.float64)
        scale = tgt_std / sd
        bias = seg.classifier.bias.double() * scale + (tgt_mean - m * scale)
        out[f"seg{SEG_SEED}_scale"] = scale.float().numpy()
        out[f"seg{SEG_SEED}_bias"] = bias.float().numpy()

        emb = nets.make_embedding(EMB_SEED, calibrated=False)
        h = emb.sincnet(x)
        bn_i = 0
        for layer in emb.tdnns:
            if isinstance(layer, torch.nn.BatchNorm1d):
                mean, var = h.mean(dim=(0, 2)), h.var(dim=(0, 2), unbiased=False)
                layer.running_mean.copy_(mean)
                layer.running_var.copy_(var)
                out[f"emb{EMB_SEED}_bn{bn_i}_mean"] = mean.numpy().copy()
                out[f"emb{EMB_SEED}_bn{bn_i}_var"] = var.numpy().copy()
                bn_i += 1
            h = layer(h)
        w = torch.rand(128, 293, 3, generator=torch.Generator().manual_seed(5)) ** 3
        e = emb.forward_dedup(x, w)
        out[f"emb{EMB_SEED}_bias"] = (emb.embedding.bias - e.mean(dim=(0, 1))).numpy().copy()
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "diart_b200", "synth_calib.npz")
    np.savez(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})
    # report what the calibrated nets look like
    with torch.no_grad():
        seg = nets.make_segmentation(SEG_SEED)
        s = seg(x)
        print("frac active", (s.amax(1) >= 0.6).float().mean(0), "frac long", (s.mean(1) >= 0.3).float().mean(0))
        emb = nets.make_embedding(EMB_SEED)
        probs = torch.softmax(10 * s, dim=-1)
        wts = (s ** 3) * (probs ** 3)
        wts[wts < 1e-8] = 1e-8
        e = emb.forward_dedup(x, wts)
        en = e / e.norm(dim=-1, keepdim=True)
        d = 1 - en.reshape(-1, 512) @ en.reshape(-1, 512).T
        print("cosine distance quantiles", np.quantile(d.numpy(), [0.01, 0.25, 0.5, 0.75, 0.99]))


if __name__ == "__main__":
    main()
