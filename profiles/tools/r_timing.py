"""Cycle split of the tcgen05 backward row kernel (library built with -DTTG_R_TIMING, CTA 3, lane 0 of every
warp): where a tile's time goes in each role."""
import ctypes as C
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "falcon-ttdforgnns_b200"))
import torch
import _ttg
_ttg.LIB_PATH = os.path.join(ROOT, "falcon-ttdforgnns_b200", "lib", "libttg_timing.so")
import tt_embeddings as te
lib = _ttg.lib()
p, q, rr, N, D, nnz = [125, 140, 140], [4, 5, 5], [1, 16, 16, 1], 2449029, 100, 262144
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(1)
cores = [(torch.randn(1, p[t], rr[t] * q[t] * rr[t + 1], generator=g) / N ** 0.25).to(dev) for t in range(3)]
idx = torch.randperm(N, generator=g)[:nnz].to(dev)
row = torch.arange(nnz, device=dev); tb = torch.zeros_like(row)
dO = ((torch.rand(1, nnz, D, generator=g) - 0.5) * 0.2).to(dev)
def step():
    te.tt_forward(1000, 1, nnz, D, p, q, rr, None, nnz, idx, row, tb, cores)
    te.tt_dense_backward(1000, D, p, q, rr, None, nnz, idx, row, tb, dO, cores)
for _ in range(3): step()
torch.cuda.synchronize()
buf = (C.c_longlong * 16)()
lib.ttg_r_timing(buf, 1)
step(); torch.cuda.synchronize()
lib.ttg_r_timing(buf, 0)
names = ["w:cp.async wait", "w:barrier rows", "w:prefetch issue", "w:fill X'", "w:fill dO^T", "w:fill core0 op",
         "w:wait_st+arrive", "w:wait d_full", "w:G0 ld+sts", "w:barrier", "w:G0 rmw d0", "w:wait s_full",
         "w:S1 extract", "w:barrier end", "mma:waits", "mma:issue"]
tot_w = sum(buf[i] for i in range(14))
print("worker lanes summed over 4 warps (cycles, one launch, CTA 3):")
for i, n in enumerate(names):
    print("  %-18s %10d  %5.1f%%" % (n, buf[i], 100.0 * buf[i] / (tot_w if i < 14 else (buf[14] + buf[15]))))
print("  worker total / 4 warps = %d cycles; mma warp total = %d" % (tot_w // 4, buf[14] + buf[15]))
