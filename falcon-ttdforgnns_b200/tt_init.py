"""TT-core initialisers of the reference's drivers (tt_utils.py:117-201), restated on torch so that they run on
the device that will hold the cores (the reference runs them in numpy on the host and copies).

    cores, ranks = tt_matrix_decomp(matrix, tt_ranks, tt_p_shapes, tt_q_shapes)   # TT-SVD of a full table
    cores = get_ortho(tt_ranks, tt_p_shapes, tt_q_shapes)                          # orthonormal slices

Core layout is the module's: core t is [1, p_t, r_t * q_t * r_{t+1}], row i_t the row-major matrix
[r_t][q_t][r_{t+1}] (FBTT/tt_embeddings_ops.py:519-545), so `module.tt_cores[t].data.copy_(cores[t])` is how
gnn_model.py:127-139 installs an "eigen" initialisation.  Host-side utilities, not part of the hot path.
"""
from typing import List, Sequence, Tuple

import torch


def tt_matrix_decomp(matrix: torch.Tensor, tt_ranks: Sequence[int], tt_p_shapes: Sequence[int],
                     tt_q_shapes: Sequence[int]) -> Tuple[List[torch.Tensor], List[int]]:
    """TT-SVD of `matrix` [prod(p), prod(q)] into three cores (tt_utils.py:157-201): the table is viewed as a
    3-way tensor over the merged modes (p_t q_t), unfolded mode by mode, and every unfolding truncated to the
    requested rank by an SVD; the singular values travel to the right."""
    p, q = [int(x) for x in tt_p_shapes], [int(x) for x in tt_q_shapes]
    assert len(p) == 3 and len(q) == 3 and len(tt_ranks) == 4, "three cores, ranks [1, r1, r2, 1]"
    m = torch.as_tensor(matrix)
    assert m.shape == (p[0] * p[1] * p[2], q[0] * q[1] * q[2])
    work = m.to(torch.float64)                     # the SVDs in double, the cores in fp32
    temp = work.reshape(p + q).permute(0, 3, 1, 4, 2, 5).reshape([p[i] * q[i] for i in range(3)])
    dims = list(temp.shape)
    ranks = [1, 1, 1, 1]
    cores = []
    for i in range(2):
        rows = ranks[i] * dims[i]
        temp = temp.reshape(rows, -1)
        cols = temp.shape[1]
        ranks[i + 1] = 1 if int(tt_ranks[i + 1]) == 1 else min(int(tt_ranks[i + 1]), cols, rows)
        u, s, vh = torch.linalg.svd(temp, full_matrices=False)
        u, s, vh = u[:, :ranks[i + 1]], s[:ranks[i + 1]], vh[:ranks[i + 1]]
        core = u.reshape(ranks[i], p[i], q[i], ranks[i + 1]).permute(1, 0, 2, 3).reshape(1, p[i], -1)
        cores.append(core.to(torch.float32).contiguous())
        temp = s[:, None] * vh
    core = temp.reshape(ranks[2], p[2], q[2], 1).permute(1, 0, 2, 3).reshape(1, p[2], -1)
    cores.append(core.to(torch.float32).contiguous())
    return cores, ranks


def get_ortho(tt_ranks: Sequence[int], tt_p_shapes: Sequence[int], tt_q_shapes: Sequence[int],
              generator: torch.Generator = None, device="cpu") -> List[torch.Tensor]:
    """Orthonormal initialisation (tt_utils.py:117-155): for every (r_t, q_t) slice of core t one row of the Q
    factor of a random square matrix, normalised, reshaped to [p_t, r_{t+1}]."""
    p, q, r = [int(x) for x in tt_p_shapes], [int(x) for x in tt_q_shapes], [int(x) for x in tt_ranks]
    rank = r[1]
    cores = []
    for t in range(3):
        n = p[t] * (rank if t < 2 else 1)
        m = torch.randn(n, n, generator=generator, dtype=torch.float32)
        qm, _ = torch.linalg.qr(m)
        v = torch.zeros(r[t], p[t], q[t], r[t + 1], dtype=torch.float32)
        k = 0
        for i in range(r[t]):
            for j in range(q[t]):
                row = qm[k] / torch.linalg.norm(qm[k])
                v[i, :, j, :] = row.reshape(p[t], r[t + 1])
                k += 1
        cores.append(v.permute(1, 0, 2, 3).reshape(1, p[t], -1).contiguous().to(device))
    return cores
