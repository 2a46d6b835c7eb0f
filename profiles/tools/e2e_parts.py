"""Device time of the parts of the end-to-end step (CUDA events, 20 repetitions each)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "falcon-ttdforgnns_b200")]
import pipeline  # noqa: E402
from FBTT.tt_embeddings_ops import OptimType, TTEmbeddingBag  # noqa: E402

dev = torch.device("cuda", 0)
p, q, ranks, N, D, nnz = [125, 140, 140], [4, 5, 5], [16, 16], 2449029, 100, 262144
m = TTEmbeddingBag(N, D, ranks, p, q, optimizer=OptimType.SGD, learning_rate=0.01, sparse=True,
                   use_cache=False, weight_dist="normal")
g = torch.Generator().manual_seed(0)
idx_h = torch.randperm(N, generator=g)[:nnz].contiguous().pin_memory()
off_h = torch.arange(nnz + 1, dtype=torch.int64).pin_memory()
idx, off = idx_h.to(dev), off_h.to(dev)
target = (torch.rand(nnz, D, generator=g) - 0.5).to(dev) * 0.1


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def step():
    loss = (m(idx, off) * target).sum()
    loss.backward()
    return loss


gs = pipeline.GraphedStep(step, dev)
print("graph replay, whole module step      %.1f us" % timed(gs))
out = m(idx, off).detach()
print("forward only (no_grad)               %.1f us" % timed(lambda: torch.no_grad()(lambda: m(idx, off))()))
print("loss forward (mul + sum)             %.1f us" % timed(lambda: (out * target).sum()))
o2 = out.clone().requires_grad_(True)
l2 = (o2 * target).sum()
print("loss backward (autograd of mul/sum)  %.1f us" % timed(lambda: torch.autograd.grad(l2, o2, retain_graph=True)))
idx_d, off_d = torch.empty_like(idx), torch.empty_like(off)
print("h2d of indices + offsets (4.2 MB)    %.1f us" % timed(lambda: (idx_d.copy_(idx_h, non_blocking=True), off_d.copy_(off_h, non_blocking=True))))
gfwd = pipeline.GraphedStep(lambda: torch.no_grad()(lambda: m(idx, off))(), dev)
print("graph replay, forward only           %.1f us" % timed(gfwd))

tflat = target.view(-1)


def step_dot(i=idx, o=off):
    loss = torch.dot(m(i, o).view(-1), tflat)
    loss.backward()
    return loss


gd = pipeline.GraphedStep(step_dot, dev)
print("graph replay, module step, dot loss  %.1f us" % timed(gd))
pipe = pipeline.HostBatchPipeline(dev, 2)
slots = []
for _ in range(2):
    pipe.put(idx_h, off_h)
for _ in range(2):
    slots.append(pipe.get())
    pipe.release()
gsl = [pipeline.GraphedStep(lambda a=a, b=b: step_dot(a, b), dev) for a, b in slots]
rd = pipeline.DeferredScalars(dev, 1)
state = {"i": 0}
pipe.put(idx_h, off_h)


def piped(copy=True, read=True):
    i = state["i"]
    state["i"] = i + 1
    if copy:
        pipe.get()
        pipe.put(idx_h, off_h)
    loss = gsl[i % 2]()
    if copy:
        pipe.release()
    if read:
        rd.push(loss)


print("pipelined: graphs only               %.1f us" % timed(lambda: piped(False, False), 40))
print("pipelined: graphs + loss read-back   %.1f us" % timed(lambda: piped(False, True), 40))
state["i"] = 0
print("pipelined: graphs + h2d + read-back  %.1f us" % timed(lambda: piped(True, True), 40))
state["i"] = 0
print("pipelined: graphs + h2d              %.1f us" % timed(lambda: piped(True, False), 40))
