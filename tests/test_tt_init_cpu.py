"""TT-core initialisers (tt_init.py) against the reference's own numpy code where /root/reference exists
(tt_utils.py:117-201, imported unchanged) and by construction elsewhere: the TT-SVD of a table of exact TT rank
reproduces the table."""
import os
import sys

import numpy as np
import pytest
import torch

import tt_init
from FBTT.tt_embeddings_ops import tt_matrix_to_full

P, Q, R = [6, 5, 4], [2, 3, 2], [1, 4, 3, 1]


def _table(seed):
    g = torch.Generator().manual_seed(seed)
    cores = [torch.randn(1, P[t], R[t] * Q[t] * R[t + 1], generator=g) for t in range(3)]
    full = tt_matrix_to_full(P, Q, R, [c[0] for c in cores], [1, 0, 2, 3])
    return cores, full


def test_tt_svd_reproduces_a_table_of_that_rank():
    _, full = _table(1)
    cores, ranks = tt_init.tt_matrix_decomp(full, R, P, Q)
    assert ranks == R
    back = tt_matrix_to_full(P, Q, ranks, [c[0] for c in cores], [1, 0, 2, 3])
    assert float((back - full).abs().max() / full.abs().max()) < 1e-5
    # truncation: lower ranks give the best approximation the unfoldings allow, never an error
    cores2, ranks2 = tt_init.tt_matrix_decomp(full, [1, 2, 2, 1], P, Q)
    assert ranks2 == [1, 2, 2, 1] and cores2[1].shape == (1, P[1], 2 * Q[1] * 2)


def test_tt_svd_matches_the_reference_code():
    ref = "/root/reference"
    if not os.path.isdir(ref):
        pytest.skip("the reference tree is not on this machine")
    sys.path.append(ref)
    try:
        import tt_utils as ref_utils
    finally:
        sys.path.remove(ref)
    _, full = _table(2)
    want, wranks = ref_utils.tt_matrix_decomp(full.numpy().astype(np.float64), R, P, Q)
    got, ranks = tt_init.tt_matrix_decomp(full, R, P, Q)
    assert list(wranks) == ranks
    # singular vectors are determined up to sign: compare the reconstructed tables
    a = tt_matrix_to_full(P, Q, ranks, [c[0] for c in got], [1, 0, 2, 3])
    b = tt_matrix_to_full(P, Q, ranks, [c[0].float() for c in want], [1, 0, 2, 3])
    assert float((a - b).abs().max() / b.abs().max()) < 1e-5


def test_ortho_slices_are_unit_vectors():
    P = [6, 5, 8]          # the last core draws r2 * q2 rows of a p2 x p2 orthogonal matrix: p2 >= r2 * q2
    cores = tt_init.get_ortho([1, 4, 4, 1], P, Q, generator=torch.Generator().manual_seed(3))
    for t, c in enumerate(cores):
        r0, r1 = [1, 4, 4, 1][t], [1, 4, 4, 1][t + 1]
        v = c[0].reshape(P[t], r0, Q[t], r1)
        norms = torch.linalg.norm(v.permute(1, 2, 0, 3).reshape(r0 * Q[t], -1), dim=1)
        assert torch.allclose(norms, torch.ones_like(norms), atol=1e-5)
