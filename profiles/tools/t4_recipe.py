"""The reference's final GCN / GAT recipe (run_script.sh:501-541: p = 50,60,60,60 q = 4,2,4,4 ranks 16,16,16 at
ogbn-arxiv size): full-table forward and fused-SGD backward, this library's default path (first two cores merged,
T = 3 kernels) against its shape-generic kernels (TTG_FLAG_FORCE_GENERIC)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "falcon-ttdforgnns_b200")]
import torch
import tt_embeddings as te
dev = "cuda:0"
p, q, rr, N = [50, 60, 60, 60], [4, 2, 4, 4], [1, 16, 16, 16, 1], 169343
D = 128
nnz = N
g = torch.Generator().manual_seed(1)
cores = [(torch.randn(1, p[t], rr[t] * q[t] * rr[t + 1], generator=g) / N ** 0.2).to(dev) for t in range(4)]
idx = torch.arange(N, device=dev)
row = torch.arange(nnz, device=dev)
tb = torch.zeros_like(row)
dO = (torch.rand(1, nnz, D, generator=g) * 0.1).to(dev)


def t(fn, n=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for fl, name in ((0, "merged, T = 3 kernels"), (1, "shape-generic kernels")):
    te.EXTRA_FLAGS = fl
    f = t(lambda: te.tt_forward(1000, 1, nnz, D, p, q, rr, None, nnz, idx, row, tb, cores))
    b = t(lambda: te.tt_sgd_backward(1000, D, 0.0, p, q, rr, None, nnz, idx, row, tb, dO, cores))
    print("%-24s forward %.3f ms   backward + SGD %.3f ms   (all %d rows)" % (name, f, b, nnz))
