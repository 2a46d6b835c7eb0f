"""Kernel-level breakdown of one GraphSAGE training step at products shape (torch.profiler)."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "falcon-ttdforgnns_b200")]
import sage  # noqa: E402
import sampler  # noqa: E402

dev = torch.device("cuda", 0)
N, E = 2449029, 123718280
graph = sage.synthetic_graph(N, E, dev, seed=0)
labels = torch.randint(0, 47, (N,), device=dev)
model = sage.SAGE(N, 100, 256, 47, 3, 0.5, (16, 16), (125, 140, 140), (4, 5, 5), sparse=True).to(dev)
tr = sage.Trainer(model)
smp = sampler.NeighborSampler([5, 10, 15])
seeds_all = torch.randperm(N, device=dev)[:196615]


def step(i):
    seeds = seeds_all[i * 1024:(i + 1) * 1024]
    inp, outp, blocks = smp.sample_blocks(graph, seeds, seed=i)
    return tr.step(blocks, inp, labels[outp])


for i in range(5):
    step(i)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for i in range(5, 10):
        step(i)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=70))
