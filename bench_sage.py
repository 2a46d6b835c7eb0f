#!/usr/bin/env python
"""bench_sage.py -- GraphSAGE epoch time at ogbn-products shape (second half of BASELINE.json's
metric: "SAGE epoch s @products 1/2/4/8 GPU", config 2).

    python bench_sage.py [--gpus N] [--epochs E] [--batch 1024]
    python -m torch.distributed.run --nproc-per-node N ... bench_sage.py --gpus N

Synthetic graph of the products shape (2,449,029 nodes, 123,718,280 directed edges, power-law
degrees), 196,615 training seeds, 47 classes, 3 SAGEConv('mean') layers of width 256, fanouts
[5, 10, 15], TT table p=125,140,140 q=4,5,5 ranks 16,16 (sage_dgl_partition.py, tt_utils.py:42-43).
Everything of a step runs on the device through this package: neighbour sampling and block
construction (ttg_sample_block), TT reconstruction, mean aggregation (ttg_spmm_csr_*), backward,
optimizers.  N > 1: the seeds of an epoch are split over the ranks (strong scaling), gradients are
all-reduced once per step over NCCL.  One JSON line on rank 0; epoch time is the max over ranks
of CUDA-event time, first epoch discarded as warm-up.
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "falcon-ttdforgnns_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--epochs", type=int, default=2, help="timed epochs (one more runs first)")
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--nodes", type=int, default=2449029)
    ap.add_argument("--edges", type=int, default=123718280)
    ap.add_argument("--train", type=int, default=196615)
    ap.add_argument("--hidden", type=int, default=256)
    ap.add_argument("--classes", type=int, default=47)
    ap.add_argument("--flags", type=int, default=0)
    ap.add_argument("--matmul", default="fp32", choices=["fp32", "tf32"],
                    help="precision of the dense SAGE layers (torch library GEMMs, not part of the "
                         "TT path): fp32 is torch's and the reference's default")
    args = ap.parse_args()

    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench_sage.py needs CUDA devices (the product has no CPU path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    json_out = sys.stdout
    if world > 1:
        # libraries write to fd 1 (NCCL's version banner); keep a private copy for the JSON line
        sys.stdout.flush()
        json_out = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
    import dp
    import sage
    import sampler
    import tt_embeddings as te
    te.EXTRA_FLAGS = int(args.flags)
    torch.backends.cuda.matmul.allow_tf32 = (args.matmul == "tf32")

    torch.manual_seed(0)
    graph = sage.synthetic_graph(args.nodes, args.edges, dev, seed=0)   # same graph on every rank
    labels = torch.randint(0, args.classes, (args.nodes,), device=dev,
                           generator=torch.Generator(device=dev).manual_seed(1))
    train_idx = torch.randperm(args.nodes, generator=torch.Generator().manual_seed(2))[:args.train]
    model = sage.SAGE(args.nodes, 100, args.hidden, args.classes, 3, 0.5, (16, 16),
                      (125, 140, 140), (4, 5, 5), sparse=(world == 1), learning_rate=0.01).to(dev)
    if world > 1:   # identical replicas
        for p in list(model.parameters()):
            dist.broadcast(p.data, 0)
    trainer = sage.Trainer(model, lr=0.003, world=world)
    smp = sampler.NeighborSampler([5, 10, 15])

    def run_epoch(epoch):
        perm = dp.epoch_permutation(args.train, epoch, seed=3)
        lo, hi = dp.shard_range(args.train, rank, world)
        mine = train_idx[perm[lo:hi]].to(dev)
        nsteps = (mine.numel() + args.batch - 1) // args.batch
        # every rank runs the same number of steps (collectives inside)
        if world > 1:
            t = torch.tensor([nsteps], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            nsteps = int(t.item())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        stats = {"input_nodes": 0, "edges0": 0}
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        e0.record()
        loss = None
        batches = sampler.prefetched_minibatches(
            graph, smp, lambda s: mine[(s * args.batch) % max(mine.numel(), 1):][:args.batch],
            lambda s: epoch * 100003 + s * world + rank, nsteps)
        for inp, outp, blocks in batches:      # sampled one step ahead on a side stream
            loss = trainer.step(blocks, inp, labels[outp])
            stats["input_nodes"] += inp.numel()
            stats["edges0"] += blocks[0].indices.numel()
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / 1e3, nsteps, float(loss.item()), stats

    run_epoch(0)                              # warm-up: allocator, workspaces, cuBLAS handles
    secs, losses, steps, st = [], [], 0, None
    for e in range(1, args.epochs + 1):
        s, steps, l, st = run_epoch(e)
        secs.append(s)
        losses.append(l)
    trainer.close()
    if rank == 0:
        best = min(secs)
        print(json.dumps({
            "metric": "GraphSAGE epoch seconds @ogbn-products shape", "value": best, "unit": "s",
            "n_gpus": world, "higher_is_better": False, "scaling": "strong", "epochs_timed": secs,
            "steps_per_epoch_per_rank": steps, "ms_per_step": best / steps * 1e3,
            "seeds_per_s": args.train / best, "data": "synthetic", "dtype": "f32",
            "dense_layer_matmul": args.matmul,
            "loss_last": losses[-1],
            "per_step_mean": {"layer0_input_nodes": st["input_nodes"] / steps,
                              "layer0_edges": st["edges0"] / steps},
            "config": {"workload": "GraphSAGE 3x SAGEConv(mean) hidden %d, %d classes, fanout [5,10,15], "
                                   "batch %d, %d train seeds, graph %d nodes / %d directed edges, "
                                   "TT p=125,140,140 q=4,5,5 ranks 16,16" %
                                   (args.hidden, args.classes, args.batch, args.train, args.nodes,
                                    args.edges),
                       "parallelism": "dp%d, replicated model, one NCCL all-reduce of the dense layers' flat gradient buffer + the TT cores exchanged and updated by one kernel over NVLink peer memory per step" % world
                       if world > 1 else "dp1, fused TT SGD"}}), file=json_out, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
