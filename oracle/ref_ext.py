"""Loader of the reference's own CUDA extension built by oracle/build_ref.py -- TEST / BASELINE
INFRASTRUCTURE (only tests/ and bench.py's reference legs may use it).  Returns None when the
extension was not built (e.g. a checkout without /root/reference at build time)."""
import glob
import importlib.machinery
import importlib.util
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_mod = None


def load():
    global _mod
    if _mod is not None:
        return _mod
    hits = sorted(glob.glob(os.path.join(_HERE, "_ref", "tt_embeddings*.so")))
    if not hits:
        return None
    import torch  # noqa: F401  (libtorch must be loaded before the extension)
    loader = importlib.machinery.ExtensionFileLoader("tt_embeddings", hits[0])
    spec = importlib.util.spec_from_file_location("tt_embeddings", hits[0], loader=loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    _mod = mod
    return mod


_eff = None


def load_efficient():
    """The reference's Efficient_TT extension (init_cuda, Eff_TT_forward,
    Fused_Extra_Eff_TT_backward, ...) or None."""
    global _eff
    if _eff is not None:
        return _eff
    hits = sorted(glob.glob(os.path.join(_HERE, "_ref", "efficient_tt_ref*.so")))
    if not hits:
        return None
    import torch  # noqa: F401
    loader = importlib.machinery.ExtensionFileLoader("efficient_tt_ref", hits[0])
    spec = importlib.util.spec_from_file_location("efficient_tt_ref", hits[0], loader=loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    _eff = mod
    return mod
