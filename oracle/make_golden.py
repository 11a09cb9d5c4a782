"""ORACLE helper (authoring container only): writes tests/golden/*.npz from the REFERENCE'S OWN code
(imported by oracle/ref_import.py from /root/reference/src) and from the oracle networks.

    python oracle/make_golden.py

Fixtures
  cluster_traces.npz   speaker maps + centroid digests produced by the reference's
                       OnlineSpeakerClustering on the seeded streams of oracle/synth_cluster.py
  functional_kats.npz  overlapped_speech_penalty / OverlappedSpeechPenalty(normalize) /
                       normalize_embeddings known answers from reference functional.py, blocks/embedding.py
  nets.npz             segmentation scores and unit-norm embeddings of the oracle networks, computed
                       THROUGH the reference's SpeakerSegmentation / OverlapAwareSpeakerEmbedding blocks
  net_layers.npz       per-layer fingerprints (mean, std, |max|, 32 samples) of PyanNet / XVectorSincNet / the powerset
                       PyanNet / WeSpeakerResNet34 on one seeded chunk (SURVEY.md 8(c) golden kind 1); needs no reference tree:
                       python oracle/make_golden.py --layers
"""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diart_b200 import synth  # noqa: E402
from diart_b200.core import SlidingWindow, SlidingWindowFeature  # noqa: E402
from oracle import nets, ref_import  # noqa: E402
from oracle.synth_cluster import make_stream  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
CLUSTER_CONFIGS = [  # (seed, max_speakers, sigma, delta, tau, rho, K, n_chunks)
    (0, 20, 1.2, 1.0, 0.6, 0.3, 3, 400), (1, 4, 1.2, 1.0, 0.6, 0.3, 3, 400), (2, 20, 2.5, 0.8, 0.5, 0.3, 3, 400),
    (3, 6, 3.0, 0.7, 0.6, 0.2, 3, 400), (4, 3, 1.0, 1.0, 0.6, 0.3, 3, 400), (5, 20, 0.5, 0.3, 0.6, 0.3, 3, 400),
    (6, 20, 1.5, 0.9, 0.6, 0.3, 4, 400), (7, 5, 2.0, 0.9, 0.55, 0.25, 4, 400),
]


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def fingerprint(t: torch.Tensor) -> np.ndarray:
    a = t.detach().double().reshape(-1)
    idx = torch.linspace(0, a.numel() - 1, 32).long()
    return np.concatenate([[a.mean().item(), a.std().item(), a.abs().max().item()], a[idx].numpy()])


def layer_fingerprints() -> dict:
    """{tap name: fingerprint} of the four oracle networks on window 0 of synth_audio(seed=1234)"""
    x = torch.from_numpy(synth.windows(synth.synth_audio(80000 + 8000, seed=1234), 1))[:, None, :]
    w = torch.rand(1, 293, generator=torch.Generator().manual_seed(7))
    out = {}
    with torch.no_grad():
        taps = {}
        out["seg/out"] = fingerprint(nets.make_segmentation()(x, taps))
        out.update({f"seg/{k}": fingerprint(v) for k, v in taps.items()})
        taps = {}
        emb = nets.make_embedding()
        trunk = emb.trunk(x, taps)
        out.update({f"emb/{k}": fingerprint(v) for k, v in taps.items()})
        out["emb/out_weighted"] = fingerprint(emb.embedding(emb.stats_pool(trunk, w)))
        taps = {}
        nets.make_powerset_segmentation()(x, taps)
        out["powerset/log_probabilities"] = fingerprint(taps["log_probabilities"])
        wes = nets.make_wespeaker()
        fb = wes.compute_fbank(x)
        out["wespeaker/fbank"] = fingerprint(fb)
        out["wespeaker/maps"] = fingerprint(wes.resnet.maps(fb))
        out["wespeaker/out_weighted"] = fingerprint(wes(x, w))
    return out


def main():
    if "--layers" in sys.argv:
        np.savez_compressed(os.path.join(OUT, "net_layers.npz"), **{k.replace("/", "__"): v for k, v in layer_fingerprints().items()})
        print("golden written: net_layers.npz")
        return
    ref = ref_import.load()
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    # ---- clustering traces from the reference class
    out = {"configs": np.array(CLUSTER_CONFIGS, dtype=np.float64)}
    sw = SlidingWindow(start=0, duration=5 / 293, step=5 / 293)
    for cfg in CLUSTER_CONFIGS:
        seed, M, sigma, delta, tau, rho, K, n = cfg
        seg, emb = make_stream(n, seed, K=K, sigma=sigma)
        clu = ref.clustering.OnlineSpeakerClustering(tau, rho, delta, "cosine", M)
        maps = -np.ones((n, K), dtype=np.int8)
        for i in range(n):
            m = clu.identify(SlidingWindowFeature(seg[i], sw), torch.from_numpy(emb[i]))
            for s, t in zip(*m.valid_assignments()):
                maps[i, s] = t
        out[f"maps_{seed}"] = maps
        out[f"centers_sha_{seed}"] = np.array(digest(clu.centers))
        out[f"centers_head_{seed}"] = clu.centers[:, :4].copy()
        out[f"active_{seed}"] = np.array(sorted(clu.active_centers), dtype=np.int32)
    np.savez_compressed(os.path.join(OUT, "cluster_traces.npz"), **out)
    # ---- functional KATs
    rng = np.random.default_rng(7)
    seg = rng.random((3, 40, 3)).astype(np.float32)
    seg[0, :5] = 1e-4                       # clamp branch
    seg[1, :, 1] = 1e-4                     # whole column clamps to 1e-8: min == max -> NaN -> 1e-8 under normalize
    emb = rng.standard_normal((3, 3, 16)).astype(np.float32)
    F = ref.functional
    kat = {"seg": seg, "emb": emb,
           "osp_3_10": F.overlapped_speech_penalty(torch.from_numpy(seg), 3, 10).numpy(),
           "osp_2_5": F.overlapped_speech_penalty(torch.from_numpy(seg), 2, 5).numpy(),
           "osp_2p5_7": F.overlapped_speech_penalty(torch.from_numpy(seg), 2.5, 7).numpy(),
           "osp_norm": ref.embedding.OverlappedSpeechPenalty(3, 10, normalize=True)(torch.from_numpy(seg)).numpy(),
           "normalize_1": F.normalize_embeddings(torch.from_numpy(emb), 1).numpy(),
           "normalize_2p5": F.normalize_embeddings(torch.from_numpy(emb), 2.5).numpy()}
    np.savez_compressed(os.path.join(OUT, "functional_kats.npz"), **kat)
    # ---- networks through the reference's own blocks (oracle nets behind the loader API)
    seg_net, emb_net = nets.make_segmentation(), nets.make_embedding()
    stream = synth.synth_audio(80000 + 8000 * 7, seed=1234)
    x = synth.windows(stream, 8)[:2]
    seg_block = ref.segmentation.SpeakerSegmentation(ref.models.SegmentationModel(lambda: seg_net), torch.device("cpu"))
    emb_block = ref.embedding.OverlapAwareSpeakerEmbedding(ref.models.EmbeddingModel(lambda: emb_net), 3, 10, 1,
                                                           False, torch.device("cpu"))
    batch = torch.from_numpy(x)[:, :, None]              # (batch, samples, channels) as diarization.py:177 builds it
    s = seg_block(batch)
    e = emb_block(batch, s)
    np.savez_compressed(os.path.join(OUT, "nets.npz"), seg=s.numpy(), emb=e.numpy(),
                        note=np.array("reference SpeakerSegmentation / OverlapAwareSpeakerEmbedding over oracle nets; "
                                      "audio = synth_audio(seed=1234) windows 0..1"))
    print("golden written:", os.listdir(OUT))


if __name__ == "__main__":
    main()
