"""Per-step cost of the sampler and the aggregation on the blocks of a products-shaped minibatch
(batch 1024, fanouts [5, 10, 15]): CUDA events, 20 repetitions, algorithmic bytes per SURVEY 8d
(aggregation: 4 F per edge read + 4 per edge index + 4 F per destination written + 8 per indptr)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "falcon-ttdforgnns_b200")]
import gnn_ops  # noqa: E402
import sage  # noqa: E402
import sampler  # noqa: E402

dev = torch.device("cuda", 0)
N, E = 2449029, 123718280
g = sage.synthetic_graph(N, E, dev, seed=0)
seeds_all = torch.randperm(N, device=dev)[:196615]
smp = sampler.NeighborSampler([5, 10, 15])


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


state = {"i": 0}


def sample():
    i = state["i"] = (state["i"] + 1) % 150
    return smp.sample_blocks(g, seeds_all[i * 1024:(i + 1) * 1024], seed=i)


ms = timed(lambda: sample())
inp, outp, blocks = sample()
print("sampler (3 layers, incl. the three size read-backs): %.3f ms per minibatch; blocks: %s"
      % (ms, [(b.num_src, b.num_dst, b.indices.numel()) for b in blocks]))
for l, (b, F) in enumerate(zip(blocks, (100, 256, 47))):
    x = torch.randn(b.num_src, F, device=dev)
    ms_f = timed(lambda: gnn_ops.aggregate(b, x, mean=True))
    xg = x.clone().requires_grad_(True)
    out = gnn_ops.aggregate(b, xg, mean=True)
    go = torch.randn_like(out)
    ms_b = timed(lambda: torch.autograd.grad(out, xg, go, retain_graph=True))
    ne, nd = b.indices.numel(), b.num_dst
    bytes_f = ne * (4 * F + 4) + nd * (4 * F + 8)
    print("layer %d aggregate F=%3d: fwd %.3f ms (%.0f GB/s algorithmic), bwd incl. zero-fill of dx %.3f ms"
          % (l, F, ms_f, bytes_f / ms_f / 1e6, ms_b))
