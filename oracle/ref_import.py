"""ORACLE helper (authoring container only): import the reference's OWN hot-path modules.

``/root/reference`` does not exist on the GPU box, so this module is used only by
``oracle/make_golden.py`` (run here, outputs committed under ``tests/golden/``) and by the
``-m "not gpu"`` tests that are skipped when the reference tree is absent.

``import diart`` itself fails in this image (``rx``, ``pyannote.*``, ``torchaudio.io`` ... are
missing), but the hot-path modules only need ``pyannote.core`` and
``pyannote.core.utils.distance.cdist``.  We register

* ``pyannote.core``  -> ``diart_b200.core`` (the same shim the product side uses), and
* ``pyannote.core.utils.distance.cdist`` -> ``scipy.spatial.distance.cdist`` (which is what the
  real function calls for ``metric="cosine"``),

create bare package shells for ``diart`` / ``diart.blocks`` so their ``__init__`` files are not
executed, and then ``importlib`` the unmodified files
``mapping.py functional.py features.py models.py blocks/{clustering,segmentation,embedding}.py``
straight from ``/root/reference/src``.  Nothing is copied.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REF_SRC = "/root/reference/src"


def available() -> bool:
    return os.path.isdir(os.path.join(REF_SRC, "diart"))


def load():
    """Returns a namespace with the reference modules: .mapping .functional .features .models
    .clustering .segmentation .embedding"""
    if not available():
        raise RuntimeError("reference tree not present (expected only in the authoring container)")
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if repo not in sys.path:
        sys.path.insert(0, repo)
    from diart_b200 import core

    if "pyannote.core" not in sys.modules:
        import scipy.spatial.distance as ssd

        pa = types.ModuleType("pyannote")
        pa.__path__ = []
        pc = types.ModuleType("pyannote.core")
        pc.__path__ = []
        for name in ("Segment", "SlidingWindow", "SlidingWindowFeature", "Annotation"):
            setattr(pc, name, getattr(core, name))
        pcu = types.ModuleType("pyannote.core.utils")
        pcu.__path__ = []
        pcd = types.ModuleType("pyannote.core.utils.distance")
        pcd.cdist = ssd.cdist
        sys.modules.update({"pyannote": pa, "pyannote.core": pc, "pyannote.core.utils": pcu,
                            "pyannote.core.utils.distance": pcd})
    if "diart" not in sys.modules:
        d = types.ModuleType("diart")
        d.__path__ = [os.path.join(REF_SRC, "diart")]
        b = types.ModuleType("diart.blocks")
        b.__path__ = [os.path.join(REF_SRC, "diart", "blocks")]
        sys.modules["diart"] = d
        sys.modules["diart.blocks"] = b
    ns = types.SimpleNamespace()
    ns.mapping = importlib.import_module("diart.mapping")
    ns.functional = importlib.import_module("diart.functional")
    ns.features = importlib.import_module("diart.features")
    ns.models = importlib.import_module("diart.models")
    ns.clustering = importlib.import_module("diart.blocks.clustering")
    ns.segmentation = importlib.import_module("diart.blocks.segmentation")
    ns.embedding = importlib.import_module("diart.blocks.embedding")
    return ns


if __name__ == "__main__":
    ref = load()
    print("loaded:", [m for m in vars(ref)])
