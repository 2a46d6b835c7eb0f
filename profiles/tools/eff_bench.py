"""BASELINE config 3 (Efficient_TT, partition-ordered ids, batch 2048): Eff_TT_forward +
Fused_Extra_Eff_TT_backward on the layer-0 source set of a minibatch, this library against the
reference's own extension (oracle/_ref, unmodified sources built for sm_100a), CUDA events."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "falcon-ttdforgnns_b200")]
import effi_tt_embeddings as eff  # noqa: E402
import reorder  # noqa: E402
import sage  # noqa: E402
import sampler  # noqa: E402
from oracle import ref_ext  # noqa: E402

dev = torch.device("cuda", 0)
N, E, K, D = 2449029, 123718280, 125, 100
p, q, r = [125, 140, 140], [4, 5, 5], [1, 16, 16, 1]
g0, comm = sage.synthetic_community_graph(N, E, K, 0.9, dev, seed=0)
g1, _ = reorder.reorder_graph(g0, "custom", nodes_perm=torch.sort(comm, stable=True).indices)
smp = sampler.NeighborSampler([5, 10, 15])
gen = torch.Generator(device="cpu").manual_seed(3)
batches = []
for it in range(4):
    seeds = torch.randperm(N, generator=gen)[:2048].to(dev)
    inp, _, _ = smp.sample_blocks(g1, seeds, seed=it)
    batches.append(inp.contiguous())
rows = [b.numel() for b in batches]
print("layer-0 rows per minibatch:", rows)
tg = torch.Generator().manual_seed(9)
cores0 = [(torch.rand(p[t], r[t] * q[t] * r[t + 1], generator=tg) * 0.3).to(dev) for t in range(3)]
tp, tq, tr = (torch.tensor(x).to(dev) for x in (p, q, r))
dOs = [(torch.rand(n, D, generator=tg) * 0.01).to(dev) for n in rows]
uni = [b.unique(sorted=True, return_inverse=True) for b in batches]


def run(mod, name, reps):
    cores = [c.clone() for c in cores0]
    mod.init_cuda(0, q, r, max(rows), D)

    def step(k):
        out = mod.Eff_TT_forward(rows[k], N, D, batches[k], p, q, r, tp, tq, tr, cores)
        mod.Fused_Extra_Eff_TT_backward(rows[k], N, D, 1e-4, batches[k], p, q, r, tp, tq, tr, dOs[k], cores,
                                        uni[k][0], uni[k][1])
        return out
    for k in range(4):
        step(k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        step(i % 4)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("%-34s %8.3f ms/step  %.3e rows/s" % (name, ms, np.mean(rows) / ms * 1e3))
    return ms


ours = run(eff, "this library (effi_tt_embeddings)", 20)
ref = ref_ext.load_efficient()
if ref is not None:
    theirs = run(ref, "reference Efficient_TT extension", 3)
    print("ratio %.1fx" % (theirs / ours))

# ---- the FBTT LFU cache on the same index stream (a-5): warm-up statistics over 8 minibatches,
# cache_populate, then forward + backward with the cached / uncached split
from FBTT.tt_embeddings_ops import OptimType, TTEmbeddingBag  # noqa: E402

more = []
for it in range(4, 12):
    seeds = torch.randperm(N, generator=gen)[:2048].to(dev)
    inp, _, _ = smp.sample_blocks(g1, seeds, seed=it)
    more.append(inp.contiguous())
for frac in (0.0, 0.01, 0.1):
    cache = int(N * frac)
    m = TTEmbeddingBag(N, D, [16, 16], p, q, optimizer=OptimType.SGD, learning_rate=1e-4, sparse=True,
                       use_cache=cache > 0, cache_size=cache, hashtbl_size=max(4 * cache, 1),
                       weight_dist="normal")
    offs = [torch.arange(b.numel() + 1, device=dev) for b in batches]
    if cache:
        for b in more:
            m(b, torch.arange(b.numel() + 1, device=dev))
        m.cache_populate()
    tgt = [d for d in dOs]

    def step(k):
        out = m(batches[k], offs[k])
        torch.dot(out.view(-1), tgt[k].view(-1)).backward()
    for k in range(4):
        step(k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(12):
        step(i % 4)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 12
    hit = ""
    if cache:
        # fraction of a minibatch's rows that the populated cache serves
        import tt_embeddings as te
        res = te.preprocess_indices_sync(batches[0], offs[0], 1, False, m.hashtbl, m.cache_state)
        hit = ", %.1f %% of the rows served from the cache" % (100.0 * (batches[0].numel() - res[3]) / batches[0].numel())
    print("TTEmbeddingBag fwd + loss + bwd, cache %4.1f %% of the table: %.3f ms/step%s" % (100 * frac, ms, hit))
