"""Neighbour sampling on the device: the role dgl.dataloading.NeighborSampler + DataLoader play
in the reference (graphloader.py:245-261, sage_dgl_partition.py:141-154), over the C ABI
(ttg_sample_block).  DGL is not part of this image; the semantics are restated in
csrc/sampler.cu and pinned by oracle/sampler_oracle.py.

    g = CSRGraph(indptr, indices)              # in-neighbours of every node, on the GPU
    sampler = NeighborSampler([5, 10, 15])     # fanouts, input layer first (DGL order)
    input_nodes, output_nodes, blocks = sampler.sample_blocks(g, seed_nodes, seed=epoch_step)

`blocks[l]` is a gnn_ops.Block (CSR by destination, destination nodes = first num_dst source
nodes), `input_nodes` the global ids whose features layer 0 reads.
"""
import ctypes as C
import os
from dataclasses import dataclass
from typing import List, Sequence, Tuple

import torch

import _ttg
from gnn_ops import Block

_scratch = _ttg._Workspace()


@dataclass
class CSRGraph:
    indptr: torch.Tensor     # int64 [num_nodes + 1]
    indices: torch.Tensor    # int32 [num_edges]  in-neighbours

    def __post_init__(self):
        _ttg.require_cuda(self.indptr, "indptr", torch.int64)
        _ttg.require_cuda(self.indices, "indices", torch.int32)

    @property
    def num_nodes(self) -> int:
        return self.indptr.numel() - 1

    @property
    def num_edges(self) -> int:
        return self.indices.numel()


def sample_block(g: CSRGraph, dst_nodes: torch.Tensor, fanout: int, seed: int
                 ) -> Tuple[Block, torch.Tensor]:
    """One layer: sampled in-neighbours of `dst_nodes` as a Block plus the global ids of its
    source nodes (destination nodes first).  Reads two integers back (the sizes of the block)."""
    _ttg.require_cuda(dst_nodes, "dst_nodes", torch.int64)
    dev = dst_nodes.device
    num_dst = dst_nodes.numel()
    lib = _ttg.lib()
    with _ttg.on_device(dev):
        indptr = torch.empty(num_dst + 1, dtype=torch.int64, device=dev)
        indices = torch.empty(max(num_dst * fanout, 1), dtype=torch.int32, device=dev)
        src = torch.empty(max(num_dst * (fanout + 1), 1), dtype=torch.int64, device=dev)
        counts = torch.empty(2, dtype=torch.int64, device=dev)
        nbytes = lib.ttg_sample_block_workspace_bytes(num_dst, int(fanout))
        if nbytes == 0:
            raise RuntimeError("sample_block: fanout %d out of range" % fanout)
        ws = _scratch.get(dev, nbytes)
        rc = lib.ttg_sample_block(g.num_nodes, _ttg.ptr(g.indptr), _ttg.ptr(g.indices), num_dst,
                                  _ttg.ptr(dst_nodes), int(fanout), C.c_uint64(seed & (2 ** 64 - 1)),
                                  _ttg.ptr(indptr), _ttg.ptr(indices), _ttg.ptr(src),
                                  _ttg.ptr(counts), _ttg.ptr(ws), ws.numel(), _ttg.stream_of(dev))
        _ttg.check(rc, "sample_block")
        num_edges, num_src = (int(x) for x in counts.tolist())
    return Block(indptr, indices[:num_edges], num_src, num_dst), src[:num_src]


class NeighborSampler:
    """fanouts[l] in-neighbours per destination node of layer l (input layer first)."""

    def __init__(self, fanouts: Sequence[int]):
        self.fanouts = [int(f) for f in fanouts]

    def sample_blocks(self, g: CSRGraph, seed_nodes: torch.Tensor, seed: int = 0
                      ) -> Tuple[torch.Tensor, torch.Tensor, List[Block]]:
        output_nodes = seed_nodes
        blocks: List[Block] = []
        dst = seed_nodes
        for layer in reversed(range(len(self.fanouts))):
            # a different stream of draws per layer, fixed for a given (seed, layer)
            blk, src = sample_block(g, dst, self.fanouts[layer], seed * 1000003 + layer)
            blocks.insert(0, blk)
            dst = src
        return dst, output_nodes, blocks


def prefetched_minibatches(g: CSRGraph, sampler: NeighborSampler, seeds_of_step, seed_of_step,
                           num_steps: int, depth: int = 2, thread=None):
    """Yields (input_nodes, output_nodes, blocks) for steps 0 .. num_steps - 1, each sampled on a
    side stream while the consumer's earlier steps are still running on the current stream -- what
    DGL's DataLoader does for GPU sampling (use_alternate_streams + its prefetcher thread;
    sage_dgl_partition.py:141-154 takes the defaults).  The two size read-backs per layer then wait
    for the sampler's own kernels only, not for the training step queued in front of them.

    thread=True (or TTG_SAMPLER_THREAD=1; default off): the sampling calls are issued by a worker thread
    up to `depth` minibatches ahead, so the consumer's thread does nothing but enqueue training steps.
    Measured on the GraphSAGE epoch at products shape (`profiles/r2d_sage_variance*.jsonl`,
    `profiles/r2d_sage_thread.txt`): the consumer's wait for a sample drops from 1.0-1.2 ms to 0.18 ms per
    step, but the epoch is bound by the device (4.15 ms per step against ~3.9 ms of host work) and its best
    time (0.80 s) and its run-to-run spread (epochs of 0.80 .. 1.3 s on a shared host, with either setting,
    with or without the clock poll, with either allocator mode) do not change -- hence off by default.
    The draws are a pure function of (seed, node, position), so the minibatches are the same either way.
    `seeds_of_step(s)` -> int64 seed nodes on the device, `seed_of_step(s)` -> draw seed."""
    if num_steps <= 0:
        return
    dev = g.indptr.device
    main = torch.cuda.current_stream(dev)
    # TTG_SAMPLER_PRIORITY=-1: the sampler's short kernels take the next free SMs instead of queueing behind
    # the training step's grids (measured: no effect on the epoch, `profiles/r2d_sage_priority.txt`)
    side = torch.cuda.Stream(dev, priority=int(os.environ.get("TTG_SAMPLER_PRIORITY", "0")))
    side.wait_stream(main)            # the graph and the seed tensors are ready
    if thread is None:
        thread = os.environ.get("TTG_SAMPLER_THREAD", "0") == "1"

    def produce(s):
        with torch.cuda.stream(side):
            batch = sampler.sample_blocks(g, seeds_of_step(s), seed=seed_of_step(s))
            ev = torch.cuda.Event()
            ev.record(side)
        return batch, ev

    def hand_over(item):
        (inp, outp, blocks), ev = item
        main.wait_event(ev)
        for t in [inp, outp] + [x for b in blocks for x in (b.indptr, b.indices)]:
            t.record_stream(main)     # allocated under the side stream, consumed on this one
        return inp, outp, blocks

    if not thread:
        nxt = produce(0)
        for s in range(num_steps):
            batch = hand_over(nxt)
            yield batch                   # the consumer enqueues its step ...
            if s + 1 < num_steps:
                nxt = produce(s + 1)      # ... and only then does the host wait for the next sample
        return

    import queue
    import threading
    q = queue.Queue(maxsize=max(int(depth), 1))
    stop = threading.Event()

    def put(item):
        while not stop.is_set():
            try:
                q.put(item, timeout=0.05)
                return True
            except queue.Full:
                continue
        return False

    def worker():
        try:
            torch.cuda.set_device(dev)
            for s in range(num_steps):
                if stop.is_set() or not put(produce(s)):
                    return
        except BaseException as ex:   # noqa: BLE001 -- handed to the consumer, which re-raises it
            put(ex)

    t = threading.Thread(target=worker, name="ttg-sampler", daemon=True)
    t.start()
    try:
        for s in range(num_steps):
            item = q.get()
            if isinstance(item, BaseException):
                raise item
            yield hand_over(item)
    finally:
        stop.set()
        t.join(timeout=10.0)
        main.wait_stream(side)        # whatever the worker still had in flight is done before its tensors go
