"""Where the end-to-end step (host indices -> TTEmbeddingBag.forward -> loss.backward) spends its
time: host wall-clock per phase with a device synchronize after each (so phases do not overlap;
the sum is an upper bound of the pipelined step)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "falcon-ttdforgnns_b200")]
from FBTT.tt_embeddings_ops import OptimType, TTEmbeddingBag  # noqa: E402

dev = torch.device("cuda", 0)
p, q, ranks, N, D, nnz = [125, 140, 140], [4, 5, 5], [16, 16], 2449029, 100, 262144
m = TTEmbeddingBag(N, D, ranks, p, q, optimizer=OptimType.SGD, learning_rate=0.01, sparse=True,
                   use_cache=False, weight_dist="normal")
g = torch.Generator().manual_seed(0)
idx_host = torch.randperm(N, generator=g)[:nnz].contiguous().pin_memory()
off_host = torch.arange(nnz + 1, dtype=torch.int64).pin_memory()
idx_dev = torch.empty(nnz, dtype=torch.int64, device=dev)
off_dev = torch.empty(nnz + 1, dtype=torch.int64, device=dev)
target = (torch.rand(nnz, D, generator=g) - 0.5).to(dev)
acc = {}


def phase(name, fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    a = acc.setdefault(name, [0.0, 0.0])
    a[0] += t1 - t0
    a[1] += t2 - t0
    return r


def step():
    phase("h2d", lambda: (idx_dev.copy_(idx_host, non_blocking=True), off_dev.copy_(off_host, non_blocking=True)))
    out = phase("forward", lambda: m(idx_dev, off_dev))
    loss = phase("loss", lambda: (out * target).sum())
    phase("backward", lambda: loss.backward())
    phase("item", lambda: loss.item())


for _ in range(5):
    step()
acc.clear()
n = 20
for _ in range(n):
    step()
for k, (host, tot) in acc.items():
    print("%-10s host-launch %.1f us   with sync %.1f us" % (k, host / n * 1e6, tot / n * 1e6))
# pipelined
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(n):
    idx_dev.copy_(idx_host, non_blocking=True)
    off_dev.copy_(off_host, non_blocking=True)
    out = m(idx_dev, off_dev)
    loss = (out * target).sum()
    loss.backward()
    loss.item()
t1 = time.perf_counter()
print("pipelined step %.1f us" % ((t1 - t0) / n * 1e6))
