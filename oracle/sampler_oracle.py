"""CPU restatement of csrc/sampler.cu -- TEST INFRASTRUCTURE, never imported by the product.

The reference samples with DGL 2.1 (dgl.dataloading.NeighborSampler, uniform without
replacement, then to_block; graphloader.py:245-261), which is un-vendored and absent here, so
parity against DGL itself is UNPINNED.  What is pinned is the documented semantics (all
in-neighbours when the degree is at most the fanout, else `fanout` distinct ones; destination
nodes are the first source nodes) and, bit for bit, the counter-based draw the CUDA kernel uses.
"""
import numpy as np

M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x):
    x = (np.uint64(x) + np.uint64(0x9E3779B97F4A7C15)) & M64
    x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & M64
    x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & M64
    return x ^ (x >> np.uint64(31))


def draw(seed, node, pos, n):
    with np.errstate(over="ignore"):
        inner = splitmix64((np.uint64(node) * np.uint64(0x100000001B3) + np.uint64(pos)) & M64)
        h = splitmix64(np.uint64(seed) ^ inner)
    return int((int(h >> np.uint64(32)) * int(n)) >> 32)


def sample_block(indptr, indices, dst_nodes, fanout, seed):
    """-> (blk_indptr int64, blk_indices int32 (local ids), src_nodes int64)"""
    cand = []
    for v in dst_nodes:
        lo, hi = int(indptr[v]), int(indptr[v + 1])
        d = hi - lo
        if d <= fanout:
            cand.append(indices[lo:hi].astype(np.int64))
            continue
        sel = []
        for k in range(fanout):
            j = d - fanout + k
            t = draw(seed, int(v), k, j + 1)
            if t in sel:
                t = j
            sel.append(t)
        cand.append(indices[lo + np.array(sel, dtype=np.int64)].astype(np.int64))
    counts = np.array([c.size for c in cand], dtype=np.int64)
    blk_indptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    flat = np.concatenate(cand) if cand else np.zeros(0, np.int64)
    dst_nodes = np.asarray(dst_nodes, dtype=np.int64)
    local = {int(v): i for i, v in enumerate(dst_nodes)}
    new = np.array(sorted(set(int(x) for x in flat) - set(local)), dtype=np.int64)
    for i, v in enumerate(new):
        local[int(v)] = dst_nodes.size + i
    blk_indices = np.array([local[int(x)] for x in flat], dtype=np.int32)
    return blk_indptr, blk_indices, np.concatenate([dst_nodes, new])
