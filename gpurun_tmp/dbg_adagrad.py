import sys, numpy as np, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/falcon-ttdforgnns_b200"); sys.path.insert(0, "/root/repo/tests")
import _ttg, tt_embeddings as te
from oracle import oracle as orc
p, q, r, n_emb = [55, 55, 56], [4, 4, 8], [16, 16], 169343
D = 128; rr = [1] + r + [1]; cols = [rr[t] * q[t] * rr[t + 1] for t in range(3)]
g = torch.Generator().manual_seed(17)
cores = [torch.randn(1, p[t], cols[t], generator=g) / (n_emb ** 0.25) for t in range(3)]
rng = np.random.default_rng(5); nnz = 12000
idx = rng.integers(0, n_emb, size=nnz).astype(np.int64); row = np.arange(nnz, dtype=np.int64)
dO = ((rng.random(size=(1, nnz, D)) - 0.5) * 0.2).astype(np.float32)
DEV = "cuda:0"
def rel(a, b): return float(np.abs(a - b).max() / np.abs(b).max())
for fl in (1024, 32):
    te.EXTRA_FLAGS = fl
    dev = [c.to(DEV) for c in cores]; state = [torch.zeros_like(c) for c in dev]
    want = [c.numpy().copy() for c in cores]; wstate = [np.zeros_like(c) for c in want]
    for it in range(2):
        gd = te.tt_dense_backward(1000, D, p, q, rr, None, nnz, torch.from_numpy(idx).to(DEV), torch.from_numpy(row).to(DEV), torch.zeros(nnz, dtype=torch.int64, device=DEV), torch.from_numpy(dO).to(DEV), dev)
        gw = orc.tt_backward_dense(p, q, r, want, idx, row, dO)
        print(fl, it, "grad err", [rel(a.cpu().numpy(), b) for a, b in zip(gd, gw)], "nan", [bool(torch.isnan(a).any()) for a in gd])
        te.tt_adagrad_backward(1000, D, 0.05, 1e-10, p, q, rr, None, nnz, torch.from_numpy(idx).to(DEV), torch.from_numpy(row).to(DEV), torch.zeros(nnz, dtype=torch.int64, device=DEV), torch.from_numpy(dO).to(DEV), state, dev)
        orc.apply_optimizer(p, cols, "adagrad", 0.05, 1e-10, want, wstate, gw)
        print(fl, it, "state err", [rel(state[t].cpu().numpy(), wstate[t]) for t in range(3)], "core err", [rel(dev[t].cpu().numpy(), want[t]) for t in range(3)])
        d = np.abs(dev[1].cpu().numpy() - want[1]); k = np.unravel_index(d.argmax(), d.shape)
        print("   worst core1 elem", k, dev[1].cpu().numpy()[k], want[1][k], "state", state[1].cpu().numpy()[k], wstate[1][k], "g", gw[1][k])
