"""Kernel timeline of one replayed step (CUDA graph of tt_forward + tt_sgd_backward, products shape, 262,144 rows):
start offset and duration of every kernel from CUPTI (torch.profiler), to see what the step's critical path is.
    python profiles/tools/step_timeline.py [--plan-ahead]"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "falcon-ttdforgnns_b200")]
import torch
from torch.profiler import ProfilerActivity, profile
import tt_embeddings as te

ahead = "--plan-ahead" in sys.argv
if ahead and '--share' in sys.argv:
    te.EXTRA_FLAGS = 512      # TTG_FLAG_SHARE_SMS
p, q, rr, N, D, nnz = [125, 140, 140], [4, 5, 5], [1, 16, 16, 1], 2449029, 100, 262144
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(1)
cores = [(torch.randn(1, p[t], rr[t] * q[t] * rr[t + 1], generator=g) / N ** 0.5).to(dev) for t in range(3)]
idx = [torch.randperm(N, generator=g)[:nnz].to(dev) for _ in range(4)]
row = torch.arange(nnz, device=dev); tb = torch.zeros_like(row)
dO = [((torch.rand(1, nnz, D, generator=g) - 0.5) * 0.2).to(dev) for _ in range(4)]
side = torch.cuda.Stream(dev)

def step(k):
    cur = torch.cuda.current_stream(dev)
    first = "--fork-first" in sys.argv
    if ahead and first:
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            te.tt_plan(1, nnz, p, q, rr, nnz, idx[(k + 1) % 4], row, tb, (k + 1) & 1)
    te.tt_forward(1000, 1, nnz, D, p, q, rr, None, nnz, idx[k], row, tb, cores)
    if ahead and not first:
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            te.tt_plan(1, nnz, p, q, rr, nnz, idx[(k + 1) % 4], row, tb, (k + 1) & 1)
    te.tt_sgd_backward(1000, D, 0.01, p, q, rr, None, nnz, idx[k], row, tb, dO[k], cores)
    if ahead:
        cur.wait_stream(side)

for i in range(8):
    step(i % 4)
torch.cuda.synchronize()
graphs = []
cap = torch.cuda.Stream(dev)
cap.wait_stream(torch.cuda.current_stream(dev))
with torch.cuda.stream(cap):
    for k in range(4):
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=cap):
            step(k)
        graphs.append(gr)
torch.cuda.current_stream(dev).wait_stream(cap)
torch.cuda.synchronize()
for i in range(8):
    graphs[i % 4].replay()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(8):
        graphs[i % 4].replay()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0 = None
steps = []
cur = []
for e in evs:
    nm = e.name
    if "plan_kernel" in nm and not ahead or ("Memset" in nm and ahead and False):
        pass
    cur.append(e)
# print the kernels of replays 4 and 5 relative to the first kernel of replay 4
per = len(evs) // 8
sel = evs[4 * per: 6 * per]
t0 = sel[0].time_range.start
print("plan ahead:", ahead, " kernels per step:", per, " step time from timeline: %.1f us"
      % ((evs[5 * per].time_range.start - evs[4 * per].time_range.start)))
for e in sel:
    print("%8.1f  +%6.1f  %s" % (e.time_range.start - t0, e.time_range.end - e.time_range.start, e.name[:70]))
