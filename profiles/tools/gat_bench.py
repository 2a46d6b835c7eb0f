"""GAT sparse kernels on an ogbn-arxiv-shaped full graph (169,343 nodes, 2.33 M directed edges
incl. reverse edges, 3 heads x 256 hidden as in run_script.sh / tt_utils.py:42-44): edge softmax
and attention-weighted per-head aggregation, forward and backward, CUDA events; algorithmic bytes
per SURVEY 8d (feature rows gathered per edge + outputs + scores)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "falcon-ttdforgnns_b200")]
import gnn_ops  # noqa: E402
import sage  # noqa: E402

dev = torch.device("cuda", 0)
N, E, H, F = 169343, 2332486, 3, 256
g = sage.synthetic_graph(N, E, dev, seed=0)
blk = gnn_ops.Block(g.indptr, g.indices, N, N)
ne = g.indices.numel()


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
    return e0.elapsed_time(e1) / n


score = torch.randn(ne, H, device=dev, requires_grad=True)
ft = torch.randn(N, H, F, device=dev, requires_grad=True)
a = gnn_ops.edge_softmax(blk, score)
ms = timed(lambda: gnn_ops.edge_softmax(blk, score))
print("edge_softmax fwd   %.3f ms  (%.0f GB/s: scores in, weights out)" % (ms, ne * H * 8 / ms / 1e6))
ga = torch.randn_like(a)
ms = timed(lambda: torch.autograd.grad(a, score, ga, retain_graph=True))
print("edge_softmax bwd   %.3f ms  (%.0f GB/s)" % (ms, ne * H * 12 / ms / 1e6))
ad = a.detach().requires_grad_(True)
out = gnn_ops.attention_aggregate(blk, ad, ft)
ms = timed(lambda: gnn_ops.attention_aggregate(blk, ad, ft))
bytes_f = ne * (H * F * 4 + H * 4 + 4) + N * H * F * 4
print("head_spmm fwd      %.3f ms  (%.0f GB/s algorithmic)" % (ms, bytes_f / ms / 1e6))
go = torch.randn_like(out)
ms = timed(lambda: torch.autograd.grad(out, (ad, ft), go, retain_graph=True))
print("head_spmm bwd      %.3f ms  (gather over the transposed block; %.0f GB/s algorithmic)" % (ms, (2 * bytes_f) / ms / 1e6))
gnn_ops.GATHER_BACKWARD = False
ms = timed(lambda: torch.autograd.grad(out, (ad, ft), go, retain_graph=True))
print("head_spmm bwd      %.3f ms  (atomics, incl. zero-fill of d_ft; %.0f GB/s algorithmic)" % (ms, (2 * bytes_f) / ms / 1e6))
gnn_ops.GATHER_BACKWARD = True
