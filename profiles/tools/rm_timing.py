"""Where the time of the right-grouped mma.sync backward row kernel goes, per warp (library built by
build_timing_lib.py): wait for the preceding kernel, run set-up, first tile, tile loop, wait for the slowest warp."""
import ctypes as C
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "falcon-ttdforgnns_b200"))
import numpy as np
import torch
import _ttg
_ttg.LIB_PATH = os.path.join(ROOT, "falcon-ttdforgnns_b200", "lib", "libttg_timing.so")
import tt_embeddings as te
lib = _ttg.lib()
p, q, rr, N, D = [125, 140, 140], [4, 5, 5], [1, 16, 16, 1], 2449029, 100
nnz = int(os.environ.get("NNZ", "262144"))
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(1)
cores = [(torch.randn(1, p[t], rr[t] * q[t] * rr[t + 1], generator=g) / N ** 0.25).to(dev) for t in range(3)]
idx = torch.randperm(N, generator=g)[:nnz].to(dev)
row = torch.arange(nnz, device=dev); tb = torch.zeros_like(row)
dO = ((torch.rand(1, nnz, D, generator=g) - 0.5) * 0.2).to(dev)
te.EXTRA_FLAGS = _ttg.FLAG_RIGHT
def step():
    te.tt_forward(1000, 1, nnz, D, p, q, rr, None, nnz, idx, row, tb, cores)
    te.tt_dense_backward(1000, D, p, q, rr, None, nnz, idx, row, tb, dO, cores)
for _ in range(3): step()
torch.cuda.synchronize()
ph = (C.c_ulonglong * 8)()
lib.ttg_rm_phases(ph, 1)
step()
torch.cuda.synchronize()
lib.ttg_rm_phases(ph, 0)
buf = (C.c_ulonglong * (148 * 32 * 6))()
lib.ttg_rm_marks(buf)
m = np.frombuffer(buf, dtype=np.uint64).reshape(148 * 32, 6)[:148 * 16].astype(np.int64)
t0 = m[:, 0].min()
m = (m - t0) / 1e3
names = ["entry", "after wait+staging", "run set up", "first tile ready", "loop done", "kernel end"]
print("nnz", nnz, "us since the first warp entered: min / median / max over %d warps" % len(m))
for i, n in enumerate(names):
    print("  %-20s %8.1f %8.1f %8.1f" % (n, m[:, i].min(), np.median(m[:, i]), m[:, i].max()))
loop = m[:, 4] - m[:, 3]
print("  tile loop per warp   %8.1f %8.1f %8.1f" % (loop.min(), np.median(loop), loop.max()))
pn = ["bookkeeping + row requests", "wait for staged rows", "G0 products", "next group's operand", "S1 products",
      "d_core0 adds", "S1 store + loop end"]
tot = sum(ph[i] for i in range(7))
print("tile loop of CTA 3, cycles summed over its warps (clock64 between phases; asynchronous work lands where it is waited for):")
for i, n in enumerate(pn):
    print("  %-28s %10d  %5.1f%%" % (n, ph[i], 100.0 * ph[i] / tot))
rc = (C.c_ulonglong * (1024 * 6))()
lib.ttg_rc_marks(rc)
c = np.frombuffer(rc, dtype=np.uint64).reshape(1024, 6)[:p[1]].astype(np.int64)
c = (c - c[:, 0].min()) / 1e3
print("cores kernel, us since its first CTA entered: min / median / max over %d CTAs" % len(c))
for i, n in enumerate(["entry", "operands staged", "preceding kernel done", "row counts in", "first block in", "end"]):
    print("  %-22s %8.1f %8.1f %8.1f" % (n, c[:, i].min(), np.median(c[:, i]), c[:, i].max()))
