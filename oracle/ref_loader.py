"""TEST INFRASTRUCTURE ONLY -- import the unmodified reference in THIS container.

/root/reference/sparsify_clip.py imports plotting / model packages that are not
installed here (openTSNE, umap, open_clip, sentence_transformers, matplotlib;
sparsify_clip.py:15-28).  None of them is touched by the loss hot path
(sparsify_clip.py:41-64, :110-187, :334-355), so they are replaced by inert stub
modules and the file is imported as-is.  /root/reference does not exist on the GPU
box: callers must check ``available()`` first; nothing run under ``-m gpu``,
``smoke()`` or ``bench.py`` may depend on it.
"""
import importlib
import os
import sys
import types

REFERENCE_DIR = os.environ.get("SCB200_REFERENCE_DIR", "/root/reference")
_STUBS = ["openTSNE", "umap", "open_clip", "sentence_transformers",
          "matplotlib", "matplotlib.pyplot", "wandb"]


class _Inert(types.ModuleType):
    """Module whose every attribute is a do-nothing callable/class."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        obj = type(name, (), {"__init__": lambda self, *a, **k: None,
                              "__call__": lambda self, *a, **k: None})
        setattr(self, name, obj)
        return obj


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "sparsify_clip.py"))


_cached = {}


def load():
    """Return (sparsify_clip module, uniformity module) from the reference tree."""
    if "mods" in _cached:
        return _cached["mods"]
    if not available():
        raise FileNotFoundError(f"reference not present at {REFERENCE_DIR}")
    for name in _STUBS:
        try:
            importlib.import_module(name)
        except Exception:
            sys.modules[name] = _Inert(name)
            if "." in name:
                parent, child = name.rsplit(".", 1)
                setattr(sys.modules[parent], child, sys.modules[name])
    sys.path.insert(0, REFERENCE_DIR)
    try:
        ref = importlib.import_module("sparsify_clip")
        uni = importlib.import_module("uniformity")
    finally:
        sys.path.remove(REFERENCE_DIR)
    _cached["mods"] = (ref, uni)
    return ref, uni
