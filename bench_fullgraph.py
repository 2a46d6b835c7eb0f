"""BASELINE config 4: full-graph GCN and GAT on a synthetic ogbn-arxiv-shaped graph (169,343 nodes,
2.33 M directed edges, 40 classes) with a TT node-embedding table p = 55,55,56 q = 4,4,8 ranks 16,16
(gcn_gat_partition.py: every epoch reconstructs all N rows, runs the model on the whole graph and
takes the loss on the training nodes).  Prints one JSON line per model: epoch milliseconds (CUDA
events over --epochs epochs after a warm-up), and how much of it the TT table (rows_range forward +
full backward + SGD) takes.  No DGL: gnn_ops.GCN / GAT on csrc/spmm.cu and csrc/gat.cu."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import _cublas_emulation  # noqa: E402,F401 -- before torch (see the module)

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "falcon-ttdforgnns_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def fullgraph_records(dev, models=("gcn", "gat"), epochs=10, hidden=256, layers=3, heads=3):
    """One record per model (a list of dicts): what main() prints, for bench.py's `fullgraph` sub-record."""
    import types
    args = types.SimpleNamespace(epochs=epochs, hidden=hidden, layers=layers, heads=heads, models=",".join(models))
    if not torch.cuda.is_available():
        raise RuntimeError("bench_fullgraph.py needs a CUDA device (the product has no CPU path)")
    import gnn_ops
    import sage
    from FBTT.tt_embeddings_ops import OptimType, TTEmbeddingBag
    records = []
    N, E, C, D = 169343, 2332486, 40, 128
    g = sage.synthetic_graph(N, E, dev, seed=0)
    graph = gnn_ops.Block(g.indptr, g.indices, N, N)
    labels = torch.randint(0, C, (N,), device=dev, generator=torch.Generator(device=dev).manual_seed(1))
    train_idx = torch.randperm(N, generator=torch.Generator().manual_seed(2))[:90941].to(dev)
    ids = torch.arange(N, device=dev)
    offsets = torch.arange(N + 1, device=dev)

    def timed(fn, n):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    for name in args.models.split(","):
        torch.manual_seed(0)
        emb = TTEmbeddingBag(N, D, [16, 16], [55, 55, 56], [4, 4, 8], optimizer=OptimType.SGD,
                             learning_rate=0.01, sparse=True, use_cache=False, weight_dist="normal")
        if name == "gcn":
            model = gnn_ops.GCN(D, args.hidden, C, args.layers, F.relu, 0.5, use_linear=False).to(dev)
        else:
            model = gnn_ops.GAT(D, C, args.hidden, args.layers, args.heads, F.relu, 0.5).to(dev)
        opt = torch.optim.Adam(model.parameters(), lr=0.002)
        last = {}

        def epoch():
            feat = emb(ids, offsets)                      # all N rows, as gcn_gat_partition.py:93-96
            pred = model(graph, feat)
            loss = F.cross_entropy(pred[train_idx], labels[train_idx])
            opt.zero_grad(set_to_none=True)
            loss.backward()                               # TT cores: fused SGD inside
            opt.step()
            last["loss"] = loss

        def tt_only():
            feat = emb(ids, offsets)
            torch.dot(feat.view(-1), feat.detach().view(-1)).backward()

        ms = timed(epoch, args.epochs)
        ms_tt = timed(tt_only, args.epochs)
        records.append({
            "metric": "full-graph %s epoch milliseconds @ogbn-arxiv shape" % name.upper(), "value": ms,
            "unit": "ms", "higher_is_better": False, "n_gpus": 1, "epochs_timed": args.epochs,
            "tt_forward_backward_ms_incl_dot_loss": ms_tt, "loss_last": float(last["loss"]),
            "data": "synthetic", "dtype": "f32",
            "config": {"workload": "%s, %d layers, hidden %d%s, graph %d nodes / %d directed edges, "
                                   "TT p=55,55,56 q=4,4,8 ranks 16,16, all %d rows per epoch"
                                   % (name.upper(), args.layers, args.hidden,
                                      ", %d heads" % args.heads if name == "gat" else "", N, E, N)}})
    return records


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--epochs", type=int, default=10)
    ap.add_argument("--hidden", type=int, default=256)
    ap.add_argument("--layers", type=int, default=3)
    ap.add_argument("--heads", type=int, default=3)
    ap.add_argument("--models", default="gcn,gat")
    args = ap.parse_args()
    for rec in fullgraph_records(torch.device("cuda", 0), tuple(args.models.split(",")), args.epochs, args.hidden,
                                 args.layers, args.heads):
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
