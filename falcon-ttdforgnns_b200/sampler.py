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
                           num_steps: int):
    """Yields (input_nodes, output_nodes, blocks) for steps 0 .. num_steps - 1, each sampled on a
    side stream while the consumer's previous step is still running on the current stream -- what
    DGL's DataLoader does for GPU sampling (use_alternate_streams; sage_dgl_partition.py:141-154
    takes the default).  The two size read-backs per layer then wait for the sampler's own kernels
    only, not for the training step queued in front of them, so the host keeps the training stream
    fed.  `seeds_of_step(s)` -> int64 seed nodes on the device, `seed_of_step(s)` -> draw seed."""
    if num_steps <= 0:
        return
    main = torch.cuda.current_stream(g.indptr.device)
    side = torch.cuda.Stream(g.indptr.device)
    side.wait_stream(main)            # the graph and the seed tensors are ready

    def produce(s):
        with torch.cuda.stream(side):
            batch = sampler.sample_blocks(g, seeds_of_step(s), seed=seed_of_step(s))
            ev = torch.cuda.Event()
            ev.record(side)
        return batch, ev

    nxt = produce(0)
    for s in range(num_steps):
        (inp, outp, blocks), ev = nxt
        main.wait_event(ev)
        for t in [inp, outp] + [x for b in blocks for x in (b.indptr, b.indices)]:
            t.record_stream(main)     # allocated under the side stream, consumed on this one
        yield inp, outp, blocks       # the consumer enqueues its step ...
        if s + 1 < num_steps:
            nxt = produce(s + 1)      # ... and only then does the host wait for the next sample
