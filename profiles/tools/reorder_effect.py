"""What the node order does to the TT lookup of a sampled minibatch: the graph as a dataset delivers it
(scrambled ids), after reorder_graph(g, 'metis', k=125) (csrc/kway_host.cu, host), after the device-side
label propagation ('grow'), and in the planted community order (what a perfect METIS-125 would find).

    python profiles/tools/reorder_effect.py [--nodes 2449029] [--edges 24000000] [--batches 12]

Per order: partition seconds, fraction of edges between parts, and -- for minibatches of 1024 seeds, fanout
[5, 10, 15], with seeds drawn uniformly and with seeds drawn from one part ("partition-aware batching") --
the input rows per step, the (i0, i1) groups and i0 slices they touch at p = (125, 140, 140), and the time
of the TT forward + fused-SGD backward on exactly those rows (one CUDA graph replay per step).  One JSON line per order."""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for _p in (ROOT, os.path.join(ROOT, "falcon-ttdforgnns_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nodes", type=int, default=2449029)
    ap.add_argument("--edges", type=int, default=24000000)
    ap.add_argument("--k", type=int, default=125)
    ap.add_argument("--batches", type=int, default=12)
    ap.add_argument("--orders", default="scrambled,grow,metis,planted")
    args = ap.parse_args()
    import pipeline
    import reorder
    import sage
    import sampler
    from FBTT.tt_embeddings_ops import OptimType, TTEmbeddingBag
    dev = torch.device("cuda", 0)
    n, k = args.nodes, args.k
    g0, comm = sage.synthetic_community_graph(n, args.edges, k, 0.9, dev, seed=0, ordered=False)
    p_shape, q_shape = [125, 140, 140], [4, 5, 5]
    torch.manual_seed(0)
    emb = TTEmbeddingBag(n, 100, [16, 16], p_shape, q_shape, optimizer=OptimType.SGD, learning_rate=0.01,
                         sparse=True, use_cache=False, weight_dist="normal")
    smp = sampler.NeighborSampler([5, 10, 15])
    dst_of_edge = torch.repeat_interleave(torch.arange(n, device=dev), g0.indptr[1:] - g0.indptr[:-1])

    def cut_fraction(labels):
        return float((labels[dst_of_edge] != labels[g0.indices.long()]).float().mean())

    for name in args.orders.split(","):
        t0 = time.perf_counter()
        if name == "scrambled":
            labels = torch.arange(n, device=dev) // ((n + k - 1) // k)      # id ranges of the raw order
            g, perm = g0, torch.arange(n, device=dev)
        else:
            if name == "metis":
                labels = reorder.kway_partition(g0, k, seed=0).long()
            elif name == "grow":
                labels = reorder.grow_partition(g0, k, seed=0).long()
            elif name == "planted":
                labels = comm
            else:
                raise SystemExit("unknown order %r" % name)
            torch.cuda.synchronize()
            t_part = time.perf_counter() - t0
            perm = reorder.partition_permutation(labels)
            g = reorder.permute_graph(g0, perm)
        torch.cuda.synchronize()
        secs = time.perf_counter() - t0
        part_of_new = labels[perm]                     # part of every node of the reordered graph
        rec = {"order": name, "partition_s": None if name in ("scrambled", "planted") else round(t_part, 2),
               "reorder_total_s": round(secs, 2), "edges_between_parts": round(cut_fraction(labels), 4),
               "largest_part": int(torch.bincount(labels, minlength=k).max()), "nodes": n,
               "directed_edges": int(g0.indices.numel()), "k": k}
        gen = torch.Generator(device=dev).manual_seed(7)
        for mode in ("uniform_seeds", "seeds_of_one_part"):
            rows, groups, slices, ms = [], [], [], []
            for b in range(args.batches):
                if mode == "uniform_seeds":
                    seeds = torch.randperm(n, device=dev, generator=gen)[:1024]
                else:
                    members = torch.nonzero(part_of_new == (b * 10) % k).flatten()
                    seeds = members[torch.randperm(members.numel(), device=dev, generator=gen)[:1024]]
                inp, _, _ = smp.sample_blocks(g, seeds, seed=b)
                offsets = torch.arange(inp.numel() + 1, device=dev)
                d_out = torch.rand(inp.numel(), 100, device=dev) * 0.1

                def step():
                    emb(inp, offsets).backward(d_out)

                gs = pipeline.GraphedStep(step, dev)      # replayed: the eager module call is host-bound (~0.25 ms)
                gs()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(5):
                    gs()
                e1.record()
                torch.cuda.synchronize()
                ms.append(e0.elapsed_time(e1) / 5)
                rows.append(inp.numel())
                groups.append(int(torch.unique(inp // p_shape[2]).numel()))
                slices.append(int(torch.unique(inp // (p_shape[1] * p_shape[2])).numel()))
            m = len(rows)
            rec[mode] = {"input_rows": sum(rows) / m, "groups_i0_i1": sum(groups) / m, "slices_i0": sum(slices) / m,
                         "rows_per_group": sum(rows) / max(sum(groups), 1), "tt_fwd_bwd_sgd_ms": sum(ms) / m,
                         "ns_per_row": 1e6 * sum(ms) / sum(rows)}
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
