"""The host-side multilevel k-way partitioner (csrc/kway_host.cu, the role of METIS behind
dgl.reorder_graph(g, 'metis', ...) at graphloader.py:370,440) through the C-ABI, no GPU.

METIS is not in this image and not under the reference tree, so parity with DGL's order is
unpinned: these are the invariants METIS' users rely on (labels in range, the balance bound,
determinism) and the cut on graphs whose optimum is known."""
import ctypes as C

import numpy as np
import pytest


def _kway(lib, indptr, indices, k, ub=1.03, seed=0, passes=0):
    import _ttg
    n = indptr.shape[0] - 1
    part = np.full(max(n, 1), -7, dtype=np.int32)
    cut = C.c_int64(-1)
    rc = lib.ttg_partition_kway(n, indptr.ctypes.data, indices.ctypes.data if indices.size else None, k, ub, seed,
                                passes, part.ctypes.data, C.byref(cut))
    assert rc == 0, _ttg.last_error()
    return part[:n], cut.value


def _csr(n, src, dst):
    """in-neighbour lists: row dst holds src"""
    order = np.argsort(dst, kind="stable")
    indptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(dst, minlength=n), out=indptr[1:])
    return indptr, np.ascontiguousarray(src[order], dtype=np.int32)


def _grid(w, h):
    ids = np.arange(w * h).reshape(h, w)
    a = np.concatenate([ids[:, :-1].ravel(), ids[:-1, :].ravel()])
    b = np.concatenate([ids[:, 1:].ravel(), ids[1:, :].ravel()])
    return _csr(w * h, np.concatenate([a, b]), np.concatenate([b, a]))


def _planted(rng, n, k, deg_in, deg_out):
    """k equal communities under scrambled ids: deg_in edges per node inside, deg_out anywhere"""
    lab = rng.permutation(n) % k
    src, dst = [], []
    for p in range(k):
        m = np.nonzero(lab == p)[0]
        src.append(rng.choice(m, size=m.shape[0] * deg_in // 2))
        dst.append(rng.choice(m, size=m.shape[0] * deg_in // 2))
    src.append(rng.integers(0, n, size=n * deg_out // 2))
    dst.append(rng.integers(0, n, size=n * deg_out // 2))
    s, d = np.concatenate(src), np.concatenate(dst)
    s, d = np.concatenate([s, d]), np.concatenate([d, s])
    indptr, indices = _csr(n, s, d)
    return indptr, indices, lab, int((lab[s] != lab[d]).sum())


def _cut(indptr, indices, part):
    dst = np.repeat(np.arange(indptr.shape[0] - 1), np.diff(indptr))
    return int((part[dst] != part[indices]).sum())


def _bound(n, k, ub):
    ideal = -(-n // k)
    return max(ideal, int(ub * ideal))


@pytest.mark.parametrize("k", [2, 4, 16, 64])
def test_grid_parts_are_balanced_and_compact(ttg_lib, k):
    indptr, indices = _grid(64, 64)
    part, cut = _kway(ttg_lib, indptr, indices, k)
    assert part.min() == 0 and part.max() == k - 1
    sizes = np.bincount(part, minlength=k)
    assert sizes.min() > 0 and sizes.max() <= _bound(4096, k, 1.03)
    assert cut == _cut(indptr, indices, part)
    # square blocks cut 2 * 64 * (sqrt(k) - 1) grid edges (x 2 directions); straight stripes for k = 2
    optimum = 2 * 64 if k == 2 else 2 * 2 * 64 * (int(round(k ** 0.5)) - 1)
    assert cut <= 2.0 * optimum, (cut, optimum)
    # a random assignment cuts (1 - 1/k) of all edges
    assert cut < 0.25 * indices.shape[0] * (1 - 1.0 / k)


def test_planted_communities_are_recovered(ttg_lib):
    rng = np.random.default_rng(0)
    n, k = 20000, 50
    indptr, indices, lab, planted_cut = _planted(rng, n, k, 12, 2)
    part, cut = _kway(ttg_lib, indptr, indices, k)
    assert np.bincount(part, minlength=k).max() <= _bound(n, k, 1.03)
    assert cut <= 1.02 * planted_cut, (cut, planted_cut)
    # the parts ARE the communities up to renaming: almost every community lies in one part
    purity = np.mean([np.bincount(part[lab == c]).max() / float((lab == c).sum()) for c in range(k)])
    assert purity > 0.97


def test_deterministic_in_the_seed(ttg_lib):
    rng = np.random.default_rng(1)
    indptr, indices, _, _ = _planted(rng, 6000, 12, 8, 3)
    a, cut_a = _kway(ttg_lib, indptr, indices, 12, seed=5)
    b, cut_b = _kway(ttg_lib, indptr, indices, 12, seed=5)
    assert np.array_equal(a, b) and cut_a == cut_b
    c, _ = _kway(ttg_lib, indptr, indices, 12, seed=6)
    assert np.bincount(c, minlength=12).max() <= _bound(6000, 12, 1.03)


def test_power_law_graph_beats_the_id_order_and_keeps_the_bound(ttg_lib):
    # preferential attachment: hubs with many leaves (what the two-hop matching is for)
    rng = np.random.default_rng(2)
    n, m = 30000, 4
    targets = np.empty((n, m), dtype=np.int64)
    pool = list(range(m))
    for v in range(m, n):
        pick = rng.integers(0, len(pool), size=m)
        targets[v] = [pool[i] for i in pick]
        pool.extend(targets[v].tolist())
        pool.extend([v] * m)
    src = np.repeat(np.arange(m, n), m)
    dst = targets[m:].ravel()
    indptr, indices = _csr(n, np.concatenate([src, dst]), np.concatenate([dst, src]))
    k = 25
    part, cut = _kway(ttg_lib, indptr, indices, k, ub=1.05)
    sizes = np.bincount(part, minlength=k)
    assert sizes.min() > 0 and sizes.max() <= _bound(n, k, 1.05)
    blocks = np.arange(n) // (n // k)           # contiguous id ranges = the graph as it comes
    assert cut < 0.8 * _cut(indptr, indices, blocks)


def test_directed_input_self_loops_duplicates_and_isolated_nodes(ttg_lib):
    # two directed 4-cliques joined by one edge, node 8 isolated, a self loop and a duplicate edge
    edges = [(a, b) for a in range(4) for b in range(4) if a < b] + \
            [(a, b) for a in range(4, 8) for b in range(4, 8) if a < b] + [(3, 4), (2, 2), (0, 1)]
    src = np.array([e[0] for e in edges])
    dst = np.array([e[1] for e in edges])
    indptr, indices = _csr(9, src, dst)
    part, cut = _kway(ttg_lib, indptr, indices, 2, ub=1.2)
    assert len(set(part[:4])) == 1 and len(set(part[4:8])) == 1 and part[0] != part[4]
    assert cut == 1
    assert np.bincount(part, minlength=2).max() <= 5


def test_edge_cases(ttg_lib):
    indptr, indices = _grid(8, 8)
    part, cut = _kway(ttg_lib, indptr, indices, 1)
    assert not part.any() and cut == 0
    part, _ = _kway(ttg_lib, indptr, indices, 64)          # one node per part
    assert sorted(part.tolist()) == list(range(64))
    empty = np.zeros(6, dtype=np.int64)                     # five nodes, no edges
    part, cut = _kway(ttg_lib, empty, np.zeros(0, dtype=np.int32), 3)
    assert cut == 0 and np.bincount(part, minlength=3).max() <= 2
    part, cut = _kway(ttg_lib, np.zeros(1, dtype=np.int64), np.zeros(0, dtype=np.int32), 1)   # no nodes
    assert part.size == 0 and cut == 0


def test_bad_arguments_fail_loudly(ttg_lib):
    import _ttg
    indptr, indices = _grid(4, 4)
    part = np.empty(16, dtype=np.int32)

    def call(ip, ix, k, ub=1.03):
        return ttg_lib.ttg_partition_kway(ip.shape[0] - 1, ip.ctypes.data, ix.ctypes.data, k, ub, 0, 0,
                                          part.ctypes.data, None)
    assert call(indptr, indices, 0) != 0 and "k=0" in _ttg.last_error()
    assert call(indptr, indices, 17) != 0
    assert call(indptr, indices, 2, ub=0.5) != 0 and "ubfactor" in _ttg.last_error()
    bad = indices.copy()
    bad[3] = 16
    assert call(indptr, bad, 2) != 0 and "out of range" in _ttg.last_error()
    dec = indptr.copy()
    dec[5] = dec[4] - 1
    assert call(dec, indices, 2) != 0


def test_partition_does_not_depend_on_the_number_of_host_threads(ttg_lib, monkeypatch):
    rng = np.random.default_rng(3)
    indptr, indices, _, _ = _planted(rng, 40000, 20, 10, 3)
    parts = []
    for threads in ("1", "4"):
        monkeypatch.setenv("TTG_KWAY_THREADS", threads)
        parts.append(_kway(ttg_lib, indptr, indices, 20, seed=2))
    assert np.array_equal(parts[0][0], parts[1][0]) and parts[0][1] == parts[1][1]
