"""TEST INFRASTRUCTURE ONLY -- never imported by the product.

CPU restatement of the graph relabelling dgl.reorder_graph performs for the reference
(graphloader.py:370, 432, 440, 449: node_subgraph under nodes_perm): new node i is old node
perm[i]; its in-neighbour list is the old one, in the old order, with every id replaced by its new
id.  Parity for csrc/reorder.cu is bit-exact.  DGL itself is not installed in this image, so the
restatement is pinned by its defining property instead (tests/test_reorder_cpu.py): the edge
multiset is preserved under the relabelling, and applying perm then its inverse is the identity.
"""
import numpy as np


def permute_csr(indptr, indices, perm):
    indptr = np.asarray(indptr, dtype=np.int64)
    indices = np.asarray(indices, dtype=np.int32)
    perm = np.asarray(perm, dtype=np.int64)
    n = indptr.shape[0] - 1
    inv = np.empty(n, dtype=np.int64)
    inv[perm] = np.arange(n, dtype=np.int64)
    deg = indptr[perm + 1] - indptr[perm]
    new_indptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(deg, out=new_indptr[1:])
    new_indices = np.empty(indices.shape[0], dtype=np.int32)
    for i in range(n):                       # small cases only
        old = perm[i]
        new_indices[new_indptr[i]:new_indptr[i + 1]] = inv[indices[indptr[old]:indptr[old + 1]]]
    return new_indptr, new_indices, inv


def edge_multiset(indptr, indices):
    """sorted array of (dst, src) pairs"""
    indptr = np.asarray(indptr, dtype=np.int64)
    dst = np.repeat(np.arange(indptr.shape[0] - 1, dtype=np.int64), np.diff(indptr))
    e = np.stack([dst, np.asarray(indices, dtype=np.int64)], axis=1)
    return e[np.lexsort((e[:, 1], e[:, 0]))]
