"""CPU tests of the boundary: the C-ABI library loads without a GPU and exports exactly what
include/ttg_b200.h declares; the Python surface mirrors the reference's names and argument
orders; argument validation fails loudly before any CUDA call."""
import ctypes
import inspect
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "ttg_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ttg_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(ttg_lib):
    import _ttg
    names = _declared_symbols()
    assert len(names) >= 20
    raw = ctypes.CDLL(_ttg.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), "libttg_b200.so does not export %s" % n
    assert sorted(_ttg.SIGNATURES) == names, "ctypes table and header disagree"


def test_library_is_sm100a_only():
    import _ttg
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    out = subprocess.run(["cuobjdump", "--list-elf", _ttg.LIB_PATH], capture_output=True,
                         text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_validation_errors_are_reported_without_a_gpu(ttg_lib):
    import _ttg
    lib = ttg_lib
    bad_dim = _ttg.make_shape([5, 5], [3, 3], [4])          # D = 9, not a multiple of 4
    rc = lib.ttg_tt_forward(ctypes.byref(bad_dim), 4, 0, None, None, None, None, None, None, 0, 0,
                            None)
    assert rc == -1 and b"multiple of 4" in lib.ttg_last_error()
    bad_rank = _ttg.Shape()
    bad_rank.T = 3
    bad_rank.num_tables = 1
    for t in range(3):
        bad_rank.p[t], bad_rank.q[t] = 4, 4
    bad_rank.r[0], bad_rank.r[1], bad_rank.r[2], bad_rank.r[3] = 2, 4, 4, 1
    rc = lib.ttg_tt_workspace_bytes(ctypes.byref(bad_rank), 8, 8)
    assert rc == 0                                            # size query refuses bad shapes
    rc = lib.ttg_cache_forward(4, 10, None, None, None, None, None)
    assert rc == -1 and b"multiple of 4" in lib.ttg_last_error()
    rc = lib.ttg_spmm_csr_fwd(4, 0, None, None, None, 1, None, None, None)
    assert rc == -1
    n_tt = ctypes.c_int32(-7)
    rc = lib.ttg_preprocess_indices(0, 1, None, None, 1, 1, 0, None, None, None, None, None, None,
                                    None, ctypes.byref(n_tt), None, 0, None)
    assert rc == 0 and n_tt.value == 0


def test_workspace_query_is_deterministic(ttg_lib):
    import _ttg
    s = _ttg.make_shape([125, 140, 140], [4, 5, 5], [16, 16])
    a = ttg_lib.ttg_tt_workspace_bytes(ctypes.byref(s), 1000, 1000)
    b = ttg_lib.ttg_tt_workspace_bytes(ctypes.byref(s), 1000, 1000)
    c = ttg_lib.ttg_tt_workspace_bytes(ctypes.byref(s), 1000, 2000)
    assert a == b and c > a > 0


def test_ops_refuse_cpu_tensors(ttg_lib):
    import tt_embeddings
    idx = torch.arange(4)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        tt_embeddings.tt_forward(1000, 1, 4, 8, [2, 2], [2, 4], [1, 2, 1], torch.tensor([2, 1]), 4,
                                 idx, idx, idx, [torch.zeros(1, 2, 4), torch.zeros(1, 2, 8)])
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        tt_embeddings.update_cache_state(idx, torch.zeros(8, dtype=torch.int64),
                                         torch.zeros(8, dtype=torch.int64))


# ---- Python surface --------------------------------------------------------------------------
REF_OPS = {  # FBTT/tt_embeddings.cpp:13-129 argument names, in order
    "tt_forward": ["batch_count", "num_tables", "B", "D", "tt_p_shapes", "tt_q_shapes", "tt_ranks",
                   "L", "nnz", "indices", "rowidx", "tableidx", "tt_cores"],
    "tt_dense_backward": ["batch_count", "D", "tt_p_shapes", "tt_q_shapes", "tt_ranks", "L", "nnz",
                          "indices", "rowidx", "tableidx", "d_output", "tt_cores"],
    "tt_sgd_backward": ["batch_count", "D", "learning_rate", "tt_p_shapes", "tt_q_shapes",
                        "tt_ranks", "L", "nnz", "indices", "rowidx", "tableidx", "d_output",
                        "tt_cores"],
    "tt_adagrad_backward": ["batch_count", "D", "learning_rate", "eps", "tt_p_shapes",
                            "tt_q_shapes", "tt_ranks", "L", "nnz", "indices", "rowidx", "tableidx",
                            "d_output", "optimizer_state", "tt_cores"],
    "update_cache_state": ["indices", "hashtbl", "cache_freq"],
    "cache_populate": ["num_embeddings", "tt_p_shapes", "tt_q_shapes", "tt_ranks", "tt_cores", "L",
                       "hashtbl", "cache_freq", "cache_state", "cache_weight"],
    "preprocess_indices_sync": ["colidx", "offsets", "num_tables", "warmup", "hashtbl",
                                "cache_state"],
    "cache_forward": ["B", "nnz", "cache_locations", "rowidx", "cache_weight", "output"],
    "cache_backward_sgd": ["nnz", "grad_output", "cache_locations", "rowidx", "learning_rate",
                           "cache_weight"],
    "cache_backward_dense": ["nnz", "grad_output", "cache_locations", "rowidx", "learning_rate",
                             "cache_weight"],
    "cache_backward_rowwise_adagrad_approx": ["nnz", "grad_output", "cache_locations", "rowidx",
                                              "learning_rate", "eps", "cache_optimizer_state",
                                              "cache_weight"],
}


def test_tt_embeddings_module_has_the_reference_ops():
    import tt_embeddings
    assert sorted(tt_embeddings.__all__) == sorted(REF_OPS)
    for name, args in REF_OPS.items():
        assert list(inspect.signature(getattr(tt_embeddings, name)).parameters) == args


def test_efficient_tt_module_has_the_reference_ops():
    import efficient_tt_table as m
    import effi_tt_embeddings as m2
    for name in ["init_cuda", "Eff_TT_forward", "Eff_TT_backward", "Fused_Eff_TT_backward",
                 "Fused_Extra_Eff_TT_backward"]:
        assert callable(getattr(m, name)) and getattr(m2, name) is getattr(m, name)
    assert list(inspect.signature(m.Eff_TT_forward).parameters) == [
        "batch_size", "table_length", "feature_dim", "index", "tt_p_shapes", "tt_q_shapes",
        "tt_ranks", "tensor_p_shape", "tensor_q_shape", "tensor_tt_ranks", "tt_cores"]
    assert list(inspect.signature(m.Fused_Extra_Eff_TT_backward).parameters)[:13] == [
        "batch_size", "table_length", "feature_dim", "learning_rate", "indices", "tt_p_shapes",
        "tt_q_shapes", "tt_ranks", "tensor_p_shape", "tensor_q_shape", "tensor_tt_ranks",
        "d_output", "tt_cores"]


def test_module_surface_matches_reference():
    from FBTT import tt_embeddings_ops as ops
    for name in ["OptimType", "BufferList", "tt_matrix_to_full", "TTLookupFunction",
                 "suggested_tt_shapes", "TableBatchedTTEmbeddingBag", "TTEmbeddingBag"]:
        assert hasattr(ops, name)
    sig = inspect.signature(ops.TTEmbeddingBag.__init__)
    assert list(sig.parameters)[1:] == [
        "num_embeddings", "embedding_dim", "tt_ranks", "tt_p_shapes", "tt_q_shapes", "optimizer",
        "learning_rate", "eps", "sparse", "use_cache", "cache_size", "hashtbl_size", "weight_dist",
        "enforce_embedding_dim", "batch_count"]
    d = {k: v.default for k, v in sig.parameters.items()}
    assert d["learning_rate"] == 0.1 and d["eps"] == 1e-10 and d["sparse"] is True
    assert d["use_cache"] is True and d["weight_dist"] == "approx-normal" and d["batch_count"] == 1000
    sig_t = inspect.signature(ops.TableBatchedTTEmbeddingBag.__init__)
    assert list(sig_t.parameters)[1] == "num_tables" and sig_t.parameters["use_cache"].default is False
    assert list(inspect.signature(ops.TTEmbeddingBag.forward).parameters) == [
        "self", "indices", "offsets", "warmup"]
    assert [m.value for m in ops.OptimType] == [
        "sgd", "exact_sgd", "lamb", "adam", "exact_adagrad", "exact_row_wise_adagrad", "lars_sgd",
        "partial_row_wise_adam", "partial_row_wise_lamb"]
    from Efficient_TT.efficient_tt import Eff_TTEmbedding
    assert list(inspect.signature(Eff_TTEmbedding.__init__).parameters)[1:] == [
        "num_embeddings", "embedding_dim", "tt_ranks", "tt_p_shapes", "tt_q_shapes", "optimizer",
        "learning_rate", "weight_dist", "device", "batch_size"]
    assert list(inspect.signature(Eff_TTEmbedding.forward).parameters) == [
        "self", "indices", "offsets", "unique", "inverse"]


def test_constructing_without_cuda_fails_loudly():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from FBTT.tt_embeddings_ops import TTEmbeddingBag
    with pytest.raises(AssertionError):
        TTEmbeddingBag(100, 8, [2, 2], [5, 5, 4], [2, 2, 2])


def test_tt_matrix_to_full_matches_reference_rows(golden):
    from FBTT.tt_embeddings_ops import tt_matrix_to_full
    c = golden["rowlayout"]
    cores = [torch.from_numpy(c["core%d" % t]) for t in range(3)]
    W = tt_matrix_to_full([14, 14, 14], [4, 4, 8], [16, 16], cores, [1, 0, 2, 3])
    assert W.shape == (2744, 128) and W.dtype == torch.float32
    np.testing.assert_allclose(W[torch.from_numpy(c["rows"])].numpy(), c["values"], rtol=1e-5,
                               atol=1e-6)


def test_tt_matrix_to_full_reproduces_reference_forward(golden):
    from FBTT.tt_embeddings_ops import tt_matrix_to_full
    for name in ["products_small_bags", "two_cores", "four_cores"]:
        c = golden[name]
        T = len(c["p"])
        cores = [torch.from_numpy(c["core%d" % t]) for t in range(T)]
        W = tt_matrix_to_full(list(c["p"]), list(c["q"]), list(c["ranks"]), cores, [1, 0, 2, 3])
        out = torch.nn.functional.embedding_bag(torch.from_numpy(c["indices"]), W,
                                                torch.from_numpy(c["offsets"]), mode="sum",
                                                include_last_offset=True)
        np.testing.assert_allclose(out.numpy(), c["out"], rtol=1e-5, atol=1e-6)


def test_suggested_tt_shapes_known_answers():
    import json
    from FBTT.tt_embeddings_ops import suggested_tt_shapes
    kat = json.load(open(os.path.join(ROOT, "tests", "golden", "hash_kat.json")))
    ref = kat["suggested_tt_shapes"]
    assert suggested_tt_shapes(2708, 3) == ref["2708_3"] == [10, 15, 20]
    assert suggested_tt_shapes(128, 3, allow_round_up=False) == ref["128_3_noround"] == [4, 4, 8]
    assert suggested_tt_shapes(100, 3, allow_round_up=False) == ref["100_3_noround"]
    assert suggested_tt_shapes(169343, 3) == ref["169343_3"]
    assert suggested_tt_shapes(2449029, 3) == [125, 140, 140]     # run_script.sh recipes


def test_buffer_list():
    from FBTT.tt_embeddings_ops import BufferList
    bl = BufferList("state", [torch.zeros(2), torch.ones(3)])
    bl.append(torch.full((1,), 7.0))
    assert len(bl) == 3 and bl[1].sum() == 3 and [b.numel() for b in bl] == [2, 3, 1]
    assert sorted(k for k, _ in bl.named_buffers()) == ["state0", "state1", "state2"]
