"""Drop-in check of the module surface: the reference's own gnn_model.py, imported UNCHANGED from
/root/reference with this package first on sys.path, binds `from FBTT.tt_embeddings_ops import TTEmbeddingBag`
(gnn_model.py:17) to this repo's class, and the keyword arguments it constructs the class with
(gnn_model.py:113-125) are accepted by it.  DGL / OGB are not in the image: stub modules stand in for them
(gnn_model.py only names their classes at import time).  Skipped where /root/reference does not exist (the GPU
box); nothing is executed on a device.
"""
import importlib
import inspect
import os
import sys
import types

import pytest

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "falcon-ttdforgnns_b200")


class _Anything:
    """Stands for any class / function of a stubbed package."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        return _Anything()


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        if name == "DGLError":
            return type("DGLError", (Exception,), {})
        if self.__name__.startswith("dgl.nn"):       # layers end up in nn.ModuleList: real (empty) Modules
            import torch

            def init(obj, *a, **k):
                torch.nn.Module.__init__(obj)
            return type(name, (torch.nn.Module,), {"__init__": init})
        return type(name, (_Anything,), {})


STUBS = ["dgl", "dgl.nn", "dgl.nn.pytorch", "dgl.nn.pytorch.utils", "dgl.nn.functional", "dgl.function",
         "dgl._ffi", "dgl._ffi.base", "dgl.utils", "dgl.data", "dgl.dataloading", "ogb", "ogb.graphproppred",
         "ogb.graphproppred.mol_encoder", "ogb.utils", "ogb.utils.features", "ogb.nodeproppred"]


@pytest.fixture
def reference_gnn_model():
    if not os.path.isdir(REF):
        pytest.skip("the reference tree is not on this machine")
    saved_path, saved_mods = list(sys.path), dict(sys.modules)
    for name in STUBS:
        m = _StubModule(name)
        m.__path__ = []
        sys.modules[name] = m
    for name in STUBS:                      # parent.child is the stub submodule, not a dummy class
        if "." in name:
            parent, child = name.rsplit(".", 1)
            setattr(sys.modules[parent], child, sys.modules[name])
    for name in [n for n in sys.modules if n == "FBTT" or n.startswith("FBTT.") or n in ("gnn_model", "tt_utils")]:
        del sys.modules[name]
    sys.path[:0] = [PKG, REF]          # this package shadows the reference's FBTT / Efficient_TT
    try:
        yield importlib.import_module("gnn_model")
    finally:
        sys.path[:] = saved_path
        for name in list(sys.modules):
            if name not in saved_mods:
                del sys.modules[name]
        sys.modules.update(saved_mods)


def test_reference_gnn_model_binds_to_this_package(reference_gnn_model):
    gm = reference_gnn_model
    assert os.path.realpath(gm.__file__) == os.path.realpath(os.path.join(REF, "gnn_model.py"))
    cls = gm.TTEmbeddingBag
    assert os.path.realpath(inspect.getsourcefile(cls)).startswith(os.path.realpath(PKG))
    # the constructor call of gnn_model.py:113-125, keyword for keyword
    sig = inspect.signature(cls.__init__)
    sig.bind(None, num_embeddings=2449029, embedding_dim=100, tt_ranks=[16, 16], tt_p_shapes=[125, 140, 140],
             tt_q_shapes=[4, 5, 5], sparse=True, use_cache=True, cache_size=24490, hashtbl_size=2449029,
             weight_dist="normal", batch_count=1000)
    # what the drivers touch afterwards (sage_dgl_partition.py:92,359-361; gnn_model.py:199-231)
    for attr in ("forward", "cache_populate", "reset_cache"):
        assert callable(getattr(cls, attr)), attr
    fwd = inspect.signature(cls.forward)
    fwd.bind(None, "indices", "offsets")
    # the reference's SAGE class itself is importable and keeps its constructor (use_tt / embed_name / cache knobs)
    params = inspect.signature(gm.SAGE.__init__).parameters
    for name in ("use_tt", "tt_rank", "p_shapes", "q_shapes", "embed_name", "use_cached", "cache_size"):
        assert name in params, name


def test_reference_model_without_tt_runs_its_constructor(reference_gnn_model):
    """use_tt=False builds only stubbed DGL layers: the reference's own constructor code runs end to end."""
    gm = reference_gnn_model
    import torch.nn.functional as F
    m = gm.SAGE(1000, 100, 16, 4, 3, F.relu, 0.5, use_tt=False, device="cpu")
    assert m.use_tt is False
