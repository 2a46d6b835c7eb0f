"""Host time per call of the module path (tiny batch, so the device is never the limit):
wall clock over 300 un-synchronised steps of forward / loss / backward / staging."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "falcon-ttdforgnns_b200")]
import pipeline  # noqa: E402
from FBTT.tt_embeddings_ops import OptimType, TTEmbeddingBag  # noqa: E402

dev = torch.device("cuda", 0)
p, q, ranks, N, D, nnz = [125, 140, 140], [4, 5, 5], [16, 16], 2449029, 100, 2048
m = TTEmbeddingBag(N, D, ranks, p, q, optimizer=OptimType.SGD, learning_rate=0.01, sparse=True,
                   use_cache=False, weight_dist="normal")
g = torch.Generator().manual_seed(0)
idx_h = torch.randperm(N, generator=g)[:nnz].contiguous().pin_memory()
off_h = torch.arange(nnz + 1, dtype=torch.int64).pin_memory()
idx, off = idx_h.to(dev), off_h.to(dev)
target = (torch.rand(nnz, D, generator=g) - 0.5).to(dev)


def wall(fn, n=300):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    return (t1 - t0) / n * 1e6


def fwd():
    with torch.no_grad():
        return m(idx, off)


def fwd_grad():
    return m(idx, off)


def fwd_loss():
    return (m(idx, off) * target).sum()


def full():
    (m(idx, off) * target).sum().backward()


pipe = pipeline.HostBatchPipeline(dev, 2)
rd = pipeline.DeferredScalars(dev, 1)
pipe.put(idx_h, off_h)


def piped():
    a, b = pipe.get()
    pipe.put(idx_h, off_h)
    loss = (m(a, b) * target).sum()
    loss.backward()
    pipe.release()
    rd.push(loss)


print("forward, no_grad            %.1f us" % wall(fwd))
print("forward, autograd           %.1f us" % wall(fwd_grad))
print("forward + loss              %.1f us" % wall(fwd_loss))
print("forward + loss + backward   %.1f us" % wall(full))
print("pipelined step              %.1f us" % wall(piped))
import cProfile, pstats
pr = cProfile.Profile()
pr.enable()
for _ in range(200):
    piped()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(28)
