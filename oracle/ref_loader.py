"""Oracle (TEST INFRASTRUCTURE): load the few reference source files that import cleanly by path.

``import pyannote.audio`` is impossible in the build container (lightning, pyannote.core,
asteroid_filterbanks ... are absent) but five leaf files only need torch/numpy/scipy/einops.  They are
executed *where they lie* under /root/reference (never copied) to (a) pin the oracle restatement and
(b) generate the committed fixtures in tests/golden/ (see tests/golden/make_golden.py).

/root/reference does not exist on the GPU box: nothing that runs there may call this module.
"""

from __future__ import annotations

import importlib.util
import os
import sys
import types

REF = os.environ.get("PYANNOTE_REFERENCE", "/root/reference")
SPA = os.path.join(REF, "src", "pyannote", "audio")


def available() -> bool:
    return os.path.isdir(SPA)


def _stub(name):
    if name not in sys.modules:
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
    return sys.modules[name]


def _load(modname, relpath):
    if modname in sys.modules and getattr(sys.modules[modname], "__file__", None):
        return sys.modules[modname]
    spec = importlib.util.spec_from_file_location(modname, os.path.join(SPA, relpath))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    return mod


def load_all():
    """Returns dict(pooling, receptive_field, resnet, powerset, vbx) of reference modules."""
    if not available():
        raise RuntimeError("reference tree not available (expected in the build container only)")
    for n in ["pyannote", "pyannote.audio", "pyannote.audio.models", "pyannote.audio.models.blocks",
              "pyannote.audio.utils", "pyannote.audio.models.embedding",
              "pyannote.audio.models.embedding.wespeaker"]:
        _stub(n)
    out = {}
    out["receptive_field"] = _load("pyannote.audio.utils.receptive_field", "utils/receptive_field.py")
    out["pooling"] = _load("pyannote.audio.models.blocks.pooling", "models/blocks/pooling.py")
    out["resnet"] = _load("pyannote.audio.models.embedding.wespeaker.resnet",
                          "models/embedding/wespeaker/resnet.py")
    out["powerset"] = _load("pyannote.audio.utils.powerset", "utils/powerset.py")
    out["vbx"] = _load("pyannote.audio.utils.vbx", "utils/vbx.py")
    return out
