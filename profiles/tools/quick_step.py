import sys, os, ctypes as C
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/falcon-ttdforgnns_b200")
import torch, numpy as np
import _ttg, tt_embeddings as te
lib = _ttg.lib()
p, q, rr, N, D = [125,140,140], [4,5,5], [1,16,16,1], 2449029, 100
nnz = int(os.environ.get("NNZ", "262144"))
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(1)
cores = [(torch.randn(1, p[t], rr[t]*q[t]*rr[t+1], generator=g) / N**0.25).to(dev) for t in range(3)]
idx = [torch.randperm(N, generator=g)[:nnz].to(dev) for _ in range(4)]
row = torch.arange(nnz, device=dev); tb = torch.zeros_like(row)
dO = [((torch.rand(1, nnz, D, generator=g) - 0.5) * 0.2).to(dev) for _ in range(4)]
def step(k):
    te.tt_forward(1000, 1, nnz, D, p, q, rr, None, nnz, idx[k], row, tb, cores)
    te.tt_sgd_backward(1000, D, 0.01, p, q, rr, None, nnz, idx[k], row, tb, dO[k], cores)
for fl in [int(x) for x in (sys.argv[1:] or ["0"])]:
    te.EXTRA_FLAGS = fl
    for i in range(5): step(i % 4)
    torch.cuda.synchronize()
    lib.ttg_profile_enable(1)
    for i in range(12): step(i % 4)
    torch.cuda.synchronize()
    out = {}
    i = 0
    while lib.ttg_profile_name(i):
        tot, cnt = C.c_double(0), C.c_int64(0)
        lib.ttg_profile_read(i, C.byref(tot), C.byref(cnt))
        if cnt.value: out[lib.ttg_profile_name(i).decode()] = round(tot.value / cnt.value * 1e3, 1)
        i += 1
    lib.ttg_profile_enable(0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20): step(i % 4)
    e1.record(); torch.cuda.synchronize()
    print("nnz", nnz, "flags", fl, "CPS", os.environ.get("TTG_R_CPS"), "step_us", round(e0.elapsed_time(e1) / 20 * 1e3, 1), out, flush=True)
