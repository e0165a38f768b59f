"""ORACLE support (test infrastructure only): import shims so the reference's own Python classes load from
/root/reference under this image's torch 2.11 / transformers 5.5.

Only oracle/make_golden.py uses this, and only in the build container: /root/reference does not exist on the GPU box.
Nothing is copied from the reference; its modules are imported in place.
"""
import os
import sys
import importlib.machinery
import types

REFERENCE_ROOT = os.environ.get("LLARK_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "m2t"))


def install():
    """Make `import m2t...` and `import jukebox.main` (the reference's file) work."""
    if not reference_available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # stubs for packages the image lacks; the hot path never calls into them
    def _mod(name):
        m = types.ModuleType(name)
        m.__spec__ = importlib.machinery.ModuleSpec(name, loader=None)
        return m

    if "braceexpand" not in sys.modules:
        m = _mod("braceexpand")
        m.braceexpand = lambda s: [s]
        sys.modules["braceexpand"] = m
    if "webdataset" not in sys.modules:
        m = _mod("webdataset")
        for name in ("WebDataset", "warn_and_continue", "split_by_node", "imagehandler"):
            setattr(m, name, lambda *a, **k: None)
        sys.modules["webdataset"] = m


def load_reference_jukebox_main():
    """Import the reference's jukebox/main.py in place (it needs `librosa` only for file decoding; a transient
    stub stands in for it and is removed again so transformers does not mistake it for the real package)."""
    import importlib.util

    def _no_load(*a, **k):
        raise RuntimeError("librosa is not installed; synthetic waveforms only")

    had = "librosa" in sys.modules
    if not had:
        m = types.ModuleType("librosa")
        m.__spec__ = importlib.machinery.ModuleSpec("librosa", loader=None)
        m.load = _no_load
        sys.modules["librosa"] = m
    try:
        spec = importlib.util.spec_from_file_location("ref_jukebox_main",
                                                      os.path.join(REFERENCE_ROOT, "jukebox", "main.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        if not had:
            sys.modules.pop("librosa", None)
    return mod
