"""Shared helpers for the parity tests."""
import numpy as np


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    denom = max(float(np.abs(b).max()), 1e-30)
    return float(np.abs(a - b).max() / denom)


def case_cores(c):
    T = len(c["p"])
    return [c["core%d" % t] for t in range(T)]


def collision_free_keys(orc, size, n, rng, lo=0, hi=None):
    """Keys whose primary hash slots are pairwise distinct and not adjacent, so that GPU
    insertion order cannot matter (the reference's insert is race-dependent under collisions)."""
    hi = hi or size
    out, used = [], set()
    for k in rng.permutation(np.arange(lo, hi))[: 8 * n]:
        s = orc.hash32(int(k), size)
        if s in used or (s + 1) % size in used or (s - 1) % size in used:
            continue
        used.add(s)
        out.append(int(k))
        if len(out) == n:
            break
    return np.array(out, dtype=np.int64)


def random_block(rng, num_src, num_dst, max_deg):
    deg = rng.integers(0, max_deg + 1, size=num_dst)
    deg[0] = 0
    indptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int64)
    indices = rng.integers(0, num_src, size=int(indptr[-1])).astype(np.int32)
    return indptr, indices
