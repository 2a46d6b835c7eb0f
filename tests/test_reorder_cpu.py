"""The relabelling oracle against its defining properties (no GPU)."""
import numpy as np

from oracle import reorder_oracle as ro


def _graph(rng, n, e):
    deg = rng.multinomial(e, np.ones(n) / n)
    indptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int64)
    return indptr, rng.integers(0, n, size=e).astype(np.int32)


def test_permutation_preserves_the_edge_multiset_and_round_trips():
    rng = np.random.default_rng(0)
    n, e = 200, 1500
    indptr, indices = _graph(rng, n, e)
    perm = rng.permutation(n).astype(np.int64)
    ip2, ix2, inv = ro.permute_csr(indptr, indices, perm)
    # edge (dst, src) of the new graph, mapped back through perm, is an edge of the old graph
    e_new = ro.edge_multiset(ip2, ix2)
    e_back = np.stack([perm[e_new[:, 0]], perm[e_new[:, 1]]], axis=1)
    e_back = e_back[np.lexsort((e_back[:, 1], e_back[:, 0]))]
    assert np.array_equal(e_back, ro.edge_multiset(indptr, indices))
    # neighbour order inside a row is the old order
    i = 17
    old = perm[i]
    assert np.array_equal(perm[ix2[ip2[i]:ip2[i + 1]]], indices[indptr[old]:indptr[old + 1]])
    # applying the inverse permutation restores the graph bit for bit
    ip3, ix3, _ = ro.permute_csr(ip2, ix2, inv)
    assert np.array_equal(ip3, indptr) and np.array_equal(ix3, indices)


def test_identity_and_empty_rows():
    indptr = np.array([0, 0, 2, 2, 3], dtype=np.int64)
    indices = np.array([3, 0, 1], dtype=np.int32)
    ip, ix, _ = ro.permute_csr(indptr, indices, np.arange(4))
    assert np.array_equal(ip, indptr) and np.array_equal(ix, indices)
    ip, ix, _ = ro.permute_csr(indptr, indices, np.array([3, 2, 1, 0]))
    assert ip.tolist() == [0, 1, 1, 3, 3] and ix.tolist() == [2, 0, 3]
