#!/usr/bin/env python
"""bench_sage.py -- GraphSAGE epoch time at ogbn-products shape (second half of BASELINE.json's
metric: "SAGE epoch s @products 1/2/4/8 GPU", configs 2 and 3).

    python bench_sage.py [--gpus N] [--epochs E] [--batch 1024] [--config 2|3]
    python -m torch.distributed.run --nproc-per-node N ... bench_sage.py --gpus N

Synthetic graph of the products shape (2,449,029 nodes, 123,718,280 directed edges, power-law
degrees), 196,615 training seeds, 47 classes, 3 SAGEConv('mean') layers of width 256, fanouts
[5, 10, 15], TT table p=125,140,140 q=4,5,5 ranks 16,16 (sage_dgl_partition.py, tt_utils.py:42-43).
Everything of a step runs on the device through this package: neighbour sampling and block
construction (ttg_sample_block), TT reconstruction, mean aggregation (ttg_spmm_csr_*), backward,
optimizers.  N > 1: the seeds of an epoch are split over the ranks (strong scaling), gradients are
all-reduced once per step over NCCL.  Epoch time is the max over ranks of CUDA-event time, the first
epoch discarded as warm-up.

config 3 (BASELINE.json): the same model on a graph whose node ids are laid out as 125 contiguous
partitions (what graphloader.py:399-454 produces with METIS-125: an edge stays inside its partition
with probability 0.9), the Efficient_TT embedding (--emb-name eff) and batch 2048.  At N > 1 the
Efficient_TT update (local fused SGD, replicas would diverge: SURVEY 3.5) is replaced by the same kernels'
dense gradients + the exchange step, so the replicas stay identical.

`sage_epoch_record` is what bench.py calls for the `sage` / `config3` records of its JSON line.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
import _cublas_emulation  # noqa: E402,F401 -- before torch: the dense layers' fp32 GEMMs on cuBLAS 12.9's BF16x9 emulation

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "falcon-ttdforgnns_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

DEFAULTS = dict(nodes=2449029, edges=123718280, train=196615, hidden=256, classes=47)


def sage_epoch_record(world, rank, dev, config=2, epochs=2, batch=None, flags=0, matmul="fp32",
                      nodes=DEFAULTS["nodes"], edges=DEFAULTS["edges"], train=DEFAULTS["train"],
                      hidden=DEFAULTS["hidden"], classes=DEFAULTS["classes"], clock_sampler=None,
                      fuse_input=False, other_steps=40):
    """One warm-up epoch + `epochs` timed ones on an initialised process group (world > 1) or alone.
    Returns the record (every rank; the timing is already the max over ranks)."""
    import torch.distributed as dist
    import dp
    import sage
    import sampler
    import tt_embeddings as te
    old_flags = te.EXTRA_FLAGS
    te.EXTRA_FLAGS = int(flags)
    old_tf32 = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = (matmul == "tf32")
    batch = batch or (2048 if config == 3 else 1024)
    torch.manual_seed(0)
    if config == 3:
        graph, _ = sage.synthetic_community_graph(nodes, edges, 125, 0.9, dev, seed=0, ordered=True)
    else:
        graph = sage.synthetic_graph(nodes, edges, dev, seed=0)   # same graph on every rank
    labels = torch.randint(0, classes, (nodes,), device=dev,
                           generator=torch.Generator(device=dev).manual_seed(1))
    train_idx = torch.randperm(nodes, generator=torch.Generator().manual_seed(2))[:train]
    eff = (config == 3 and world == 1)
    model = sage.SAGE(nodes, 100, hidden, classes, 3, 0.5, (16, 16), (125, 140, 140), (4, 5, 5),
                      sparse=(world == 1), learning_rate=0.01, embed_name="eff" if eff else "fbtt",
                      device=dev, fuse_input=fuse_input).to(dev)
    if world > 1:   # identical replicas
        for p in list(model.parameters()):
            dist.broadcast(p.data, 0)
    trainer = sage.Trainer(model, lr=0.003, world=world)
    smp = sampler.NeighborSampler([5, 10, 15])

    def run_epoch(epoch, max_steps=None):
        perm = dp.epoch_permutation(train, epoch, seed=3)
        lo, hi = dp.shard_range(train, rank, world)
        mine = train_idx[perm[lo:hi]].to(dev)
        nsteps = (mine.numel() + batch - 1) // batch
        if max_steps is not None:
            nsteps = min(nsteps, max_steps)
        if world > 1:     # every rank runs the same number of steps (collectives inside)
            t = torch.tensor([nsteps], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            nsteps = int(t.item())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        stats = {"input_nodes": 0, "edges0": 0}
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)
        e0.record()
        loss = None
        batches = sampler.prefetched_minibatches(
            graph, smp, lambda s: mine[(s * batch) % max(mine.numel(), 1):][:batch],
            lambda s: epoch * 100003 + s * world + rank, nsteps)
        for inp, outp, blocks in batches:      # sampled one step ahead on a side stream
            loss = trainer.step(blocks, inp, labels[outp])
            stats["input_nodes"] += inp.numel()
            stats["edges0"] += blocks[0].indices.numel()
        e1.record()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / 1e3, nsteps, float(loss.item()), stats

    run_epoch(0)                              # warm-up: allocator, workspaces, cuBLAS handles
    sampler_proc = clock_sampler() if clock_sampler else None
    secs, losses, steps, st = [], [], 0, None
    for e in range(1, epochs + 1):
        s, steps, l, st = run_epoch(e)
        secs.append(s)
        losses.append(l)
    clocks = sampler_proc.stop() if sampler_proc else None
    # the same steps with the other data flow of the first layer (a short run): fused = the neighbour mean taken
    # by the TT lookup itself, plain = all num_src rows reconstructed and then aggregated (gnn_model.py:199-217)
    ms_other = None
    if not eff and other_steps > 0:
        model.fuse_input = not model.fuse_input
        run_epoch(epochs + 1, max_steps=8)
        s_u, n_u, _, _ = run_epoch(epochs + 2, max_steps=other_steps)
        ms_other = s_u / n_u * 1e3
        model.fuse_input = not model.fuse_input
    # replicas must still be identical after the epochs (cores and dense layers)
    identical = None
    if world > 1:
        with torch.no_grad():
            flat = torch.cat([p.detach().reshape(-1).view(torch.int32).to(torch.int64)
                              for p in model.parameters()])
            chk = torch.stack([flat.sum(), (flat * torch.arange(1, flat.numel() + 1, device=dev)).sum()])
        allc = [torch.empty_like(chk) for _ in range(world)]
        dist.all_gather(allc, chk)
        identical = all(bool(torch.equal(allc[0], c)) for c in allc)
    failed = trainer.failed_epoch() if hasattr(trainer, "failed_epoch") else 0
    trainer.close()
    te.EXTRA_FLAGS = old_flags
    torch.backends.cuda.matmul.allow_tf32 = old_tf32
    best = min(secs)
    return {
        "metric": "GraphSAGE epoch seconds @ogbn-products shape", "value": best, "unit": "s",
        "n_gpus": world, "higher_is_better": False, "scaling": "strong", "epochs_timed": secs,
        "epoch_median_s": sorted(secs)[len(secs) // 2],
        "epoch_spread": "value = best timed epoch; single epochs take up to 1.6x as long on a shared host "
                        "(isolated stalls of 50-100 ms in the host's sampling calls, profiles/r2d_sage_variance2.jsonl)",
        "steps_per_epoch_per_rank": steps, "ms_per_step": best / steps * 1e3,
        "first_layer": ("neighbour mean taken by the TT lookup (one EmbeddingBag call, bags = destinations + their "
                        "sampled neighbours); [num_src, 100] is never written" if model.fuse_input else
                        "all num_src rows reconstructed, then aggregated (gnn_model.py:199-217)"),
        ("ms_per_step_plain_first_layer" if model.fuse_input else "ms_per_step_fused_first_layer"): ms_other,
        "seeds_per_s": train / best, "data": "synthetic", "dtype": "f32",
        "dense_layer_matmul": matmul + (" on " + _cublas_emulation.WHY + "; fp32-accurate: maximum error against fp64 "
                                        "1.6e-7 - 1.9e-7 of the largest element at these shapes, native SGEMM 5e-7 - "
                                        "8e-7 (profiles/r2d_cublas_*.txt)"
                                        if _cublas_emulation.ACTIVE and matmul == "fp32" else ""),
        "loss_last": losses[-1], "clocks": clocks,
        "replicas_bit_identical": identical, "exchange_failed": failed,
        "per_step_mean": {"layer0_input_nodes": st["input_nodes"] / steps,
                          "layer0_edges": st["edges0"] / steps},
        "note": "aggregation, sampler and block builder are pinned to this repo's own restatement of DGL 2.1 "
                "(DGL is not in the image): parity-unpinned rows a-9 / f-1; the TT part is about 10 % of a step, "
                "about half is torch's fp32 SGEMM in the dense SAGE layers",
        "config": {"workload": "BASELINE config %d: GraphSAGE 3x SAGEConv(mean) hidden %d, %d classes, fanout "
                               "[5,10,15], batch %d, %d train seeds, %s graph %d nodes / %d directed edges, "
                               "TT p=125,140,140 q=4,5,5 ranks 16,16, embedding %s" %
                               (config, hidden, classes, batch, train,
                                "125-partition-ordered" if config == 3 else "power-law", nodes, edges,
                                "Efficient_TT (fused SGD)" if eff else "FBTT TTEmbeddingBag"),
                   "parallelism": ("dp%d, replicated model, one NCCL all-reduce of the dense layers' flat gradient "
                                   "buffer + the TT cores exchanged and updated by one kernel over NVLink peer "
                                   "memory per step" % world) if world > 1 else "dp1, fused TT SGD"}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--epochs", type=int, default=2, help="timed epochs (one more runs first)")
    ap.add_argument("--batch", type=int, default=0, help="0: 1024 (config 2) / 2048 (config 3)")
    ap.add_argument("--config", type=int, default=2, choices=[2, 3])
    ap.add_argument("--nodes", type=int, default=DEFAULTS["nodes"])
    ap.add_argument("--edges", type=int, default=DEFAULTS["edges"])
    ap.add_argument("--train", type=int, default=DEFAULTS["train"])
    ap.add_argument("--hidden", type=int, default=DEFAULTS["hidden"])
    ap.add_argument("--classes", type=int, default=DEFAULTS["classes"])
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
    sampler_factory = None
    if rank == 0:
        from bench import ClockSampler
        sampler_factory = lambda: ClockSampler(local)
        if os.environ.get("TTG_NO_CLOCKS") == "1":     # experiment: does the nvidia-smi poll perturb the epochs?
            sampler_factory = None
    rec = sage_epoch_record(world, rank, dev, config=args.config, epochs=args.epochs,
                            batch=args.batch or None, flags=args.flags, matmul=args.matmul,
                            nodes=args.nodes, edges=args.edges, train=args.train, hidden=args.hidden,
                            classes=args.classes, clock_sampler=sampler_factory)
    if rank == 0:
        print(json.dumps(rec), file=json_out, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
