"""cProfile of the host side of TTEmbeddingBag.forward + backward (launch path only)."""
import cProfile
import os
import pstats
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "falcon-ttdforgnns_b200")]
from FBTT.tt_embeddings_ops import OptimType, TTEmbeddingBag  # noqa: E402

dev = torch.device("cuda", 0)
p, q, ranks, N, D, nnz = [125, 140, 140], [4, 5, 5], [16, 16], 2449029, 100, 262144
m = TTEmbeddingBag(N, D, ranks, p, q, optimizer=OptimType.SGD, learning_rate=0.01, sparse=True,
                   use_cache=False, weight_dist="normal")
g = torch.Generator().manual_seed(0)
idx = torch.randperm(N, generator=g)[:nnz].to(dev)
off = torch.arange(nnz + 1, dtype=torch.int64, device=dev)
target = (torch.rand(nnz, D, generator=g) - 0.5).to(dev)


def step():
    out = m(idx, off)
    loss = (out * target).sum()
    loss.backward()


for _ in range(10):
    step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(100):
    step()
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)
