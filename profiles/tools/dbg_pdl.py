import os, sys
import numpy as np, torch
ROOT = "/root/repo"
sys.path[:0] = [ROOT, os.path.join(ROOT, "falcon-ttdforgnns_b200")]
import tt_embeddings as te
DEV = "cuda:0"
p, q, r, n_emb = [125, 140, 140], [4, 5, 5], [1, 16, 16, 1], 2449029
g = torch.Generator().manual_seed(7)
cores0 = [(torch.randn(1, p[t], r[t] * q[t] * r[t + 1], generator=g) * (0.5 / np.sqrt(r[t]))).to(DEV) for t in range(3)]
nnz = 30000
rng = np.random.default_rng(4)
idx = torch.from_numpy(rng.integers(0, n_emb, size=nnz).astype(np.int64)).to(DEV)
row = torch.arange(nnz, device=DEV)
tb = torch.zeros(nnz, dtype=torch.int64, device=DEV)
dO = (torch.rand(1, nnz, 100, generator=torch.Generator().manual_seed(5)) * 0.1).to(DEV)
res = {}
for name, fl in (("generic", 1), ("mma", 0), ("mma2", 0), ("ffma", 16)):
    te.EXTRA_FLAGS = fl
    c = [x.clone() for x in cores0]
    for rep in range(1):
        te.tt_sgd_backward(1000, 100, 0.1, p, q, r, None, nnz, idx, row, tb, dO, c)
    res[name] = c
    gd = te.tt_dense_backward(1000, 100, p, q, r, None, nnz, idx, row, tb, dO, [x.clone() for x in cores0])
    res[name + "_dense"] = gd
for name in ("mma", "mma2", "ffma"):
    print(name, [float((a - b).abs().max() / (b - c0).abs().max()) for a, b, c0 in zip(res[name], res["generic"], cores0)],
          [float((a - b).abs().max() / b.abs().max()) for a, b in zip(res[name + "_dense"], res["generic_dense"])])
