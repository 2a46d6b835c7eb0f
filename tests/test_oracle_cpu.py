"""CPU tests: the oracle against the committed golden vectors (generated from the reference by
oracle/pin_against_reference.py), and the integer semantics of the index path."""
import json
import os

import numpy as np
import pytest

from helpers import case_cores, rel_err
from oracle import oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CASES = ["cora_r16", "cora_r16_bags", "products_small", "products_small_bags", "papers_small_r32",
         "two_cores", "four_cores"]


@pytest.mark.parametrize("name", CASES)
def test_forward_matches_reference_golden(golden, name):
    c = golden[name]
    B = c["offsets"].size - 1
    out = orc.tt_forward(list(c["p"]), list(c["q"]), list(c["ranks"]), case_cores(c), c["indices"],
                         c["rowidx"], B)[0]
    assert rel_err(out, c["out64"]) < 2e-6          # reference in float64
    assert rel_err(out, c["out"]) < 1e-5            # reference's own fp32 path


@pytest.mark.parametrize("name", CASES)
def test_backward_matches_reference_autograd(golden, name):
    c = golden[name]
    dc = orc.tt_backward_dense(list(c["p"]), list(c["q"]), list(c["ranks"]), case_cores(c),
                               c["indices"], c["rowidx"], c["d_output"][None])
    for t, g in enumerate(dc):
        assert rel_err(g, c["d_core%d" % t]) < 2e-6


def test_row_layout(golden):
    c = golden["rowlayout"]
    cores = [c["core%d" % t] for t in range(3)]
    rows = c["rows"]
    out = orc.tt_forward([14, 14, 14], [4, 4, 8], [1, 16, 16, 1], cores, rows,
                         np.arange(rows.size), rows.size)[0]
    assert rel_err(out, c["values"]) < 2e-6


def test_f32_rows_variant_agrees(golden):
    c = golden["products_small"]
    a = orc.tt_forward_f32_rows(list(c["p"]), list(c["q"]), list(c["ranks"]), case_cores(c),
                                c["indices"])
    assert rel_err(a, c["out64"]) < 1e-5


def test_hash_known_answers():
    kat = json.load(open(os.path.join(ROOT, "tests", "golden", "hash_kat.json")))
    assert len(kat["vectors"]) >= 100
    for key, size, h in kat["vectors"]:
        assert orc.hash32(key, size) == h
    # the vectors quoted in SURVEY.md 8c
    assert orc.hash32(0, 2449029) == 616976
    assert orc.hash32(2449028, 2449029) == 1779130
    assert orc.hash32((1 << 32) + 7, 2449029) == 217902
    assert orc.hash32(12345, 2708) == 2024


def test_lfu_update_counts_duplicates_and_drops_after_three_probes():
    size = 64
    keys = np.full(size, -1, dtype=np.int64)
    freq = np.zeros(size, dtype=np.int64)
    # four keys with the same primary slot: the fourth cannot be placed (MAX_PROBES = 3)
    same = []
    k = 0
    target = orc.hash32(0, size)
    while len(same) < 4:
        if orc.hash32(k, size) == target:
            same.append(k)
        k += 1
    idx = np.array(same + same[:2], dtype=np.int64)
    orc.update_cache_state(idx, keys, freq)
    for j in range(3):
        assert keys[(target + j) % size] == same[j]
    assert same[3] not in keys
    assert freq[target] == 2 and freq[(target + 1) % size] == 2 and freq[(target + 2) % size] == 1


def test_cache_populate_and_partition_semantics():
    rng = np.random.default_rng(0)
    size, cache_size = 257, 5
    keys = np.full(size, -1, dtype=np.int64)
    freq = np.zeros(size, dtype=np.int64)
    state = np.full(size, -1, dtype=np.int32)
    stream = rng.integers(0, 40, size=600).astype(np.int64)
    orc.update_cache_state(stream, keys, freq)
    keys0, freq0 = keys.copy(), freq.copy()
    sorted_keys = orc.cache_populate_index(cache_size, keys, freq, state)
    # top-k by frequency, ties in slot order (stable descending radix sort)
    order = sorted(range(size), key=lambda s: (-freq0[s], s))
    assert [keys0[s] for s in order[:cache_size]] == list(sorted_keys[:cache_size])
    cached = set(int(k) for k in sorted_keys[:cache_size])
    # everything else that was tracked is evicted
    assert set(int(k) for k in keys if k != -1) == cached
    for n, k in enumerate(sorted_keys[:cache_size]):
        slot = int(np.where(keys == k)[0][0])
        assert state[slot] == n
    # partition: TT items first in order, cached items at the back in reverse order
    col = rng.integers(0, 40, size=50).astype(np.int64)
    offsets = np.arange(51, dtype=np.int64)
    pc, pr, tb, n_tt, pl = orc.preprocess_indices(col, offsets, 1, False, keys, state)
    is_cached = np.array([int(k) in cached for k in col])
    assert n_tt == int((~is_cached).sum())
    assert list(pc[:n_tt]) == list(col[~is_cached])
    assert list(pr[:n_tt]) == list(np.arange(50)[~is_cached])
    assert list(pc[n_tt:]) == list(col[is_cached][::-1])
    assert list(pr[n_tt:]) == list(np.arange(50)[is_cached][::-1])
    rank = {int(k): n for n, k in enumerate(sorted_keys[:cache_size])}
    assert list(pl[n_tt:]) == [rank[int(k)] for k in col[is_cached][::-1]]
    assert (pl[:n_tt] == -1).all() and (tb == 0).all()
    # warm-up: untouched
    pc2, pr2, tb2, n2, pl2 = orc.preprocess_indices(col, offsets, 1, True, keys, state)
    assert n2 == 50 and pl2 is None and list(pc2) == list(col) and list(pr2) == list(range(50))


def test_rowidx_with_tables_and_empty_bags():
    offsets = np.array([0, 2, 2, 5, 6, 6, 9], dtype=np.int64)  # 2 tables x 3 bags
    col = np.arange(9, dtype=np.int64)
    _, rowidx, tableidx, n, _ = orc.preprocess_indices(col, offsets, 2, True, None, None)
    assert n == 9
    assert list(rowidx) == [0, 0, 2, 2, 2, 0, 2, 2, 2]
    assert list(tableidx) == [0, 0, 0, 0, 0, 1, 1, 1, 1]


def test_reference_sgd_launch_skips_tail_rows():
    # products shape -> core0 rows >= 64 and core2 rows >= 84 are never updated
    lim = orc.reference_sgd_rows_updated([125, 140, 140], [64, 1280, 80])
    assert lim == [64, 140, 84]
    lim = orc.reference_sgd_rows_updated([55, 55, 56], [64, 1024, 128])
    assert lim == [55, 55, 56]
    lim = orc.reference_sgd_rows_updated([481, 481, 481], [128, 4096, 256])
    assert lim == [128, 481, 256]


def test_optimizer_formulas():
    rng = np.random.default_rng(1)
    p, cols = [3, 4], [8, 12]
    cores = [rng.normal(size=(1, p[t], cols[t])).astype(np.float32) for t in range(2)]
    grads = [rng.normal(size=(1, p[t], cols[t])).astype(np.float32) for t in range(2)]
    c_sgd = [c.copy() for c in cores]
    orc.apply_optimizer(p, cols, "sgd", 0.1, 0.0, c_sgd, None, grads)
    for c0, c1, g in zip(cores, c_sgd, grads):
        np.testing.assert_allclose(c1, c0 - np.float32(0.1) * g, rtol=1e-6)
    c_ada = [c.copy() for c in cores]
    state = [np.zeros_like(c) for c in cores]
    orc.apply_optimizer(p, cols, "adagrad", 0.1, 1e-10, c_ada, state, grads)
    for c0, c1, g, s in zip(cores, c_ada, grads, state):
        np.testing.assert_allclose(s, g * g, rtol=1e-6)
        np.testing.assert_allclose(c1, c0 - 0.1 * g / (np.sqrt(g * g) + 1e-10), rtol=1e-5)


def test_eff_float_index_math_breaks_above_2_pow_24():
    p = [481, 481, 481]
    idx = np.array([0, 1, 480, 481, 2 ** 24 + 1, 111059955], dtype=np.int64)
    exact = orc.eff_split(idx, p, False)
    for n, i in enumerate(idx):
        g = int(i) // 481
        assert list(exact[n]) == [g, g // 481, g % 481, int(i) % 481]
    flt = orc.eff_split(np.arange(2 ** 24, 2 ** 24 + 4096, dtype=np.int64) * 6, p, True)
    ext = orc.eff_split(np.arange(2 ** 24, 2 ** 24 + 4096, dtype=np.int64) * 6, p, False)
    assert not np.array_equal(flt, ext)     # the reference's float path is inexact here


def test_spmm_against_dense():
    rng = np.random.default_rng(2)
    from helpers import random_block
    num_src, num_dst, F = 40, 17, 12
    indptr, indices = random_block(rng, num_src, num_dst, 6)
    x = rng.normal(size=(num_src, F)).astype(np.float32)
    A = np.zeros((num_dst, num_src))
    for v in range(num_dst):
        for e in range(indptr[v], indptr[v + 1]):
            A[v, indices[e]] += 1.0
    deg = np.maximum(indptr[1:] - indptr[:-1], 1)[:, None]
    np.testing.assert_allclose(orc.spmm_csr_fwd(indptr, indices, x, False), A @ x, rtol=1e-5,
                               atol=1e-6)
    np.testing.assert_allclose(orc.spmm_csr_fwd(indptr, indices, x, True), (A @ x) / deg,
                               rtol=1e-5, atol=1e-6)
    dout = rng.normal(size=(num_dst, F)).astype(np.float32)
    np.testing.assert_allclose(orc.spmm_csr_bwd(indptr, indices, dout, num_src, True),
                               A.T @ (dout / deg), rtol=1e-5, atol=1e-6)
