"""run_script.sh:247-264 (tt-ranks sweep, products shape, --q-shapes "5,5,4"): forward + dense backward per call at
262,144 distinct rows, plan-based kernels against the any-shape kernels (TTG_FLAG_FORCE_GENERIC)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "falcon-ttdforgnns_b200"))
import torch
import _ttg
import tt_embeddings as te

dev = "cuda:0"
p, q, N, nnz = [125, 140, 140], [5, 5, 4], 2449029, 262144
for rank in (8, 16, 32, 64):
    rr = [1, rank, rank, 1]
    D = q[0] * q[1] * q[2]
    g = torch.Generator().manual_seed(1)
    cores = [(torch.randn(1, p[t], rr[t] * q[t] * rr[t + 1], generator=g) / N ** 0.25).to(dev) for t in range(3)]
    idx = torch.randperm(N, generator=g)[:nnz].to(dev)
    row = torch.arange(nnz, device=dev)
    tb = torch.zeros_like(row)
    dO = (torch.rand(1, nnz, D, generator=g) * 0.1).to(dev)
    for fl, name in ((0, "plan-based"), (_ttg.FLAG_FORCE_GENERIC, "any-shape")):
        if rank == 64 and fl == 0:
            continue
        te.EXTRA_FLAGS = fl

        def step():
            te.tt_forward(1000, 1, nnz, D, p, q, rr, None, nnz, idx, row, tb, cores)
            te.tt_dense_backward(1000, D, p, q, rr, None, nnz, idx, row, tb, dO, cores)

        for _ in range(2):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            step()
        e1.record()
        torch.cuda.synchronize()
        print("q=5,5,4 ranks %d,%d %-10s %9.1f us per forward + dense backward" % (rank, rank, name, e0.elapsed_time(e1) * 200),
              flush=True)
te.EXTRA_FLAGS = 0
