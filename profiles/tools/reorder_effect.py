"""What node reordering buys the TT lookup (BASELINE config 3): a products-sized
graph with 125 planted communities and scrambled ids, minibatches of 2048 seeds with fanout
[5, 10, 15]; for every ordering the number of distinct TT groups (i0, i1) the layer-0 input nodes
touch and the time of the TT forward + backward + SGD step on them (CUDA events)."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "falcon-ttdforgnns_b200")]
import reorder  # noqa: E402
import sage  # noqa: E402
import sampler  # noqa: E402
from FBTT.tt_embeddings_ops import OptimType, TTEmbeddingBag  # noqa: E402

dev = torch.device("cuda", 0)
N, E, K = 2449029, 123718280, 125
p, q, ranks = [125, 140, 140], [4, 5, 5], [16, 16]
t0 = time.time()
g0, comm = sage.synthetic_community_graph(N, E, K, 0.9, dev, seed=0)
print("graph: %d nodes, %d directed edges, %d communities (%.1f s)" % (N, g0.num_edges, K, time.time() - t0))
emb = TTEmbeddingBag(N, 100, ranks, p, q, optimizer=OptimType.SGD, learning_rate=0.01, sparse=True,
                     use_cache=False, weight_dist="normal")
smp = sampler.NeighborSampler([5, 10, 15])


def measure(name, g, secs, batch):
    gen = torch.Generator(device="cpu").manual_seed(7)
    groups, rows, ms = [], [], []
    for it in range(12):
        # seeds of one minibatch come from one community-sized id range when the ids are ordered
        # (partition-aware batching, graphloader.py:358-372); uniformly otherwise
        seeds = torch.randperm(N, generator=gen)[:batch].to(dev)
        inp, _, _ = smp.sample_blocks(g, seeds, seed=it)
        groups.append(int(torch.unique(inp // p[2]).numel()))
        rows.append(inp.numel())
        off = torch.arange(inp.numel() + 1, device=dev)
        tgt = torch.rand(inp.numel(), 100, device=dev) * 0.01
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.dot(emb(inp, off).view(-1), tgt.view(-1)).backward()
            e1.record()
            torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    n = len(groups)
    print("batch %4d %-22s reorder %6.1f s | layer-0 rows %8.0f | distinct TT groups %7.0f (%.1f rows/group) | "
          "module fwd+loss+bwd %.3f ms" % (batch, name, secs, sum(rows) / n, sum(groups) / n,
                                          sum(rows) / sum(groups), sorted(ms)[n // 2]))


t0 = time.time()
g1, _ = reorder.reorder_graph(g0, "custom", nodes_perm=torch.sort(comm, stable=True).indices)
torch.cuda.synchronize()
t1 = time.time() - t0
t0 = time.time()
g2, _ = reorder.reorder_graph(g0, "grow", k=K, seed=0)
torch.cuda.synchronize()
t2 = time.time() - t0
g3, t3 = None, 0.0
if "--rcmk" in sys.argv:
    t0 = time.time()
    g3, _ = reorder.reorder_graph(g0, "rcmk")
    t3 = time.time() - t0
for batch in (2048, 32):
    measure("scrambled ids", g0, 0.0, batch)
    measure("planted communities", g1, t1, batch)
    measure("grow-125 (device)", g2, t2, batch)
    if g3 is not None:
        measure("rcmk (scipy, as DGL)", g3, t3, batch)
