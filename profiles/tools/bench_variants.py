"""Timing of the other call patterns of the hot path at products shape (BASELINE config 3):
TTEmbeddingBag with the LFU cache (warm), Eff_TTEmbedding, both on partition-local indices."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "falcon-ttdforgnns_b200")]
from Efficient_TT.efficient_tt import Eff_TTEmbedding  # noqa: E402
from FBTT.tt_embeddings_ops import OptimType, TTEmbeddingBag  # noqa: E402

dev = torch.device("cuda", 0)
p, q, ranks, N, D = [125, 140, 140], [4, 5, 5], [16, 16], 2449029, 100
nnz = 262144
rng = np.random.default_rng(0)


def partition_local_batch():
    """125 contiguous partitions; 90 % of a batch's ids fall into 8 of them (METIS-125 reorder)."""
    part = N // 125
    hot = rng.choice(125, size=8, replace=False)
    n_hot = int(nnz * 0.9)
    a = hot[rng.integers(0, 8, size=n_hot)] * part + rng.integers(0, part, size=n_hot)
    b = rng.integers(0, N, size=nnz - n_hot)
    ids = np.unique(np.concatenate([a, b]))
    rng.shuffle(ids)
    return torch.from_numpy(ids.astype(np.int64)).to(dev)


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (time.perf_counter() - t0) / n * 1e3


batches = [partition_local_batch() for _ in range(4)]
print("batch sizes", [b.numel() for b in batches], "groups", [int(torch.unique(b // 140).numel()) for b in batches])

# ---- FBTT with the LFU cache (cache = 5 % of the nodes, populated after a warm-up pass)
m = TTEmbeddingBag(N, D, ranks, p, q, optimizer=OptimType.SGD, learning_rate=0.01, sparse=True,
                   use_cache=True, cache_size=int(0.05 * N), hashtbl_size=N, weight_dist="normal")
for b in batches:
    m(b, torch.arange(b.numel() + 1, device=dev))
m.cache_populate()
state = {"i": 0}


def fbtt_cached():
    b = batches[state["i"] % 4]
    state["i"] += 1
    out = m(b, torch.arange(b.numel() + 1, device=dev))
    out.backward(torch.ones_like(out) * 1e-3)


ms, wall = timed(fbtt_cached)
print("FBTT + LFU cache   fwd+bwd: %.3f ms device, %.3f ms wall per step (%.3g rows/s)" %
      (ms, wall, batches[0].numel() / (wall * 1e-3)))

m2 = TTEmbeddingBag(N, D, ranks, p, q, optimizer=OptimType.SGD, learning_rate=0.01, sparse=True,
                    use_cache=False, weight_dist="normal")


def fbtt_plain():
    b = batches[state["i"] % 4]
    state["i"] += 1
    out = m2(b, torch.arange(b.numel() + 1, device=dev))
    out.backward(torch.ones_like(out) * 1e-3)


ms, wall = timed(fbtt_plain)
print("FBTT no cache      fwd+bwd: %.3f ms device, %.3f ms wall per step (%.3g rows/s)" %
      (ms, wall, batches[0].numel() / (wall * 1e-3)))

e = Eff_TTEmbedding(N, D, ranks, p, q, learning_rate=0.1, device=0, batch_size=300000)


def eff():
    b = batches[state["i"] % 4]
    state["i"] += 1
    out = e(b)
    out.backward(torch.ones_like(out) * 1e-3)


ms, wall = timed(eff)
print("Efficient_TT       fwd+bwd: %.3f ms device, %.3f ms wall per step (%.3g rows/s)" %
      (ms, wall, batches[0].numel() / (wall * 1e-3)))
