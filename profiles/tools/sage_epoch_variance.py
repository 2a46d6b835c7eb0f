"""Why do GraphSAGE epochs of one process differ (0.80 .. 1.28 s)?  Per epoch: device time, host time spent
waiting for the next sampled minibatch and enqueueing the step, allocator state, GPU temperature / clocks."""
import gc
import json
import os
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "falcon-ttdforgnns_b200")]
import dp  # noqa: E402
import sage  # noqa: E402
import sampler  # noqa: E402

dev = torch.device("cuda", 0)
N, E, TRAIN, BATCH = 2449029, 123718280, 196615, 1024
EPOCHS = int(os.environ.get("EPOCHS", "6"))     # TTG_SAMPLER_THREAD=0 / 1: who issues the sampling calls
if os.environ.get("NO_GC") == "1":
    gc.disable()
torch.manual_seed(0)
graph = sage.synthetic_graph(N, E, dev, seed=0)
labels = torch.randint(0, 47, (N,), device=dev)
train_idx = torch.randperm(N, generator=torch.Generator().manual_seed(2))[:TRAIN]
model = sage.SAGE(N, 100, 256, 47, 3, 0.5, (16, 16), (125, 140, 140), (4, 5, 5), sparse=True, learning_rate=0.01,
                  embed_name="fbtt", device=dev).to(dev)
trainer = sage.Trainer(model, lr=0.003, world=1)
smp = sampler.NeighborSampler([5, 10, 15])


def smi():
    try:
        out = subprocess.run(["nvidia-smi", "-i", "0", "--query-gpu=temperature.gpu,temperature.memory,clocks.sm,"
                              "clocks.mem,power.draw", "--format=csv,noheader,nounits"], capture_output=True,
                             text=True, timeout=10).stdout.strip()
        return out
    except Exception as ex:   # noqa: BLE001
        return str(ex)


for epoch in range(EPOCHS + 1):
    perm = dp.epoch_permutation(TRAIN, epoch, seed=3)
    mine = train_idx[perm].to(dev)
    nsteps = (TRAIN + BATCH - 1) // BATCH
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t_wait = t_step = 0.0
    per_step, evs = [], []
    e0.record()
    it = sampler.prefetched_minibatches(graph, smp, lambda s: mine[s * BATCH:][:BATCH],
                                        lambda s: epoch * 100003 + s, nsteps)
    w0 = time.perf_counter()
    while True:
        t0 = time.perf_counter()
        try:
            inp, outp, blocks = next(it)
        except StopIteration:
            break
        t1 = time.perf_counter()
        trainer.step(blocks, inp, labels[outp])
        t2 = time.perf_counter()
        t_wait += t1 - t0
        t_step += t2 - t1
        per_step.append((t1 - t0, t2 - t1))
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        evs.append(ev)
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - w0
    st = torch.cuda.memory_stats(dev)
    dev_ms = [e0.elapsed_time(evs[0])] + [evs[i].elapsed_time(evs[i + 1]) for i in range(len(evs) - 1)]
    srt = sorted(dev_ms)
    med = srt[len(srt) // 2]
    slow = [(i, round(d, 2), round(per_step[i][0] * 1e3, 2), round(per_step[i][1] * 1e3, 2))
            for i, d in enumerate(dev_ms) if d > 1.5 * med]
    print(json.dumps({"epoch": epoch, "device_s": e0.elapsed_time(e1) / 1e3, "wall_s": wall,
                      "host_wait_for_sample_ms_per_step": 1e3 * t_wait / nsteps,
                      "host_enqueue_step_ms_per_step": 1e3 * t_step / nsteps,
                      "device_ms_per_step_p50_p90_max": [round(med, 2), round(srt[int(len(srt) * 0.9)], 2), round(srt[-1], 2)],
                      "steps_over_1.5x_median": len(slow), "excess_ms": round(sum(d[1] - med for d in slow), 1),
                      "slow_steps_idx_devms_waitms_enqms": slow[:12],
                      "reserved_gb": st["reserved_bytes.all.current"] / 2 ** 30,
                      "segments": st["segment.all.current"], "alloc_retries": st["num_alloc_retries"],
                      "cudamalloc_calls": st["segment.all.allocated"],
                      "smi_tempgpu_tempmem_smclk_memclk_power": smi()}), flush=True)
