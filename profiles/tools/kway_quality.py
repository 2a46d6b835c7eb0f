"""Cut of ttg_partition_kway (csrc/kway_host.cu) against the planted partition of a community graph with scrambled
ids (the numpy twin of sage.synthetic_community_graph; host only, no GPU).

    python profiles/tools/kway_quality.py [nodes] [directed edges] [graph seeds] [partitioner seeds]

Round-2d finding: with the strict balance bound on every level, 2 of 6 runs at 1.2 M nodes / 60 M edges / k = 125 ended
7 % and 22 % above the planted cut (communities split in halves between full parts); with 1.3 x slack on the
coarse levels all runs recover the planted partition exactly (profiles/r2e_kway_quality.txt)."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "falcon-ttdforgnns_b200"))
import _ttg  # noqa: E402


def community(n, e, k, p_in, rng):
    half, size = e // 2, (n + k - 1) // k
    src = rng.integers(0, n, size=half)
    lo = (src // size) * size
    near = lo + (rng.random(half) * np.minimum(size, n - lo)).astype(np.int64)
    dst = np.where(rng.random(half) < p_in, near, rng.integers(0, n, size=half))
    scr = rng.permutation(n)
    s, d = scr[np.concatenate([src, dst])], scr[np.concatenate([dst, src])]
    order = np.argsort(d, kind="stable")
    indptr = np.zeros(n + 1, np.int64)
    np.cumsum(np.bincount(d, minlength=n), out=indptr[1:])
    comm = np.empty(n, np.int64)
    comm[scr] = np.arange(n) // size
    return indptr, s[order].astype(np.int32), comm


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1200000
    e = int(sys.argv[2]) if len(sys.argv) > 2 else 60000000
    graph_seeds = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    part_seeds = int(sys.argv[4]) if len(sys.argv) > 4 else 3
    k, lib = 125, _ttg.lib()
    for gs in range(graph_seeds):
        indptr, indices, comm = community(n, e, k, 0.9, np.random.default_rng(100 + gs))
        dst = np.repeat(np.arange(n), np.diff(indptr))
        planted = int((comm[dst] != comm[indices]).sum())
        for seed in range(part_seeds):
            part, cut = np.empty(n, np.int32), C.c_int64(0)
            t = time.time()
            rc = lib.ttg_partition_kway(n, indptr.ctypes.data, indices.ctypes.data, k, 1.03, seed, 0, part.ctypes.data,
                                        C.byref(cut))
            assert rc == 0, _ttg.last_error()
            sizes = np.bincount(part, minlength=k)
            print("n=%d e=%d k=%d graph seed %d partitioner seed %d: cut / planted cut = %.3f, parts %d..%d nodes, %.1f s"
                  % (n, e, k, 100 + gs, seed, cut.value / planted, sizes.min(), sizes.max(), time.time() - t), flush=True)


if __name__ == "__main__":
    main()
