"""Host <-> device plumbing around the hot path: the minibatch's index arrays go up on a copy
stream one step ahead of the kernels that consume them, and per-step scalars (loss, metric) come
back one step late, so that neither direction puts a host synchronisation between two steps.

The reference gets the same overlap from DGL's DataLoader (sage_dgl_partition.py:141-154: the
loader hands over batches that are already on `device` while the previous step still runs) and
pays one `loss.item()` per logging interval (:253-262).  Nothing here computes: torch is used for
pinned memory, streams and events only.
"""
from collections import deque
from typing import List, Optional, Sequence, Tuple

import torch


class HostBatchPipeline:
    """Ring of `depth` device-side staging slots fed from pinned host tensors on a copy stream.

        pipe.put(indices_host, offsets_host)        # enqueue the copies of a later step
        indices, offsets = pipe.get()               # oldest staged batch; the current stream waits
        ...forward / backward on indices, offsets...
        pipe.release()                              # this step no longer reads its slot

    A slot is overwritten only after the step that read it has been released (event on the
    compute stream that the copy stream waits for), so the index tensors a backward still needs
    are never clobbered by a prefetch."""

    def __init__(self, device: torch.device, depth: int = 2, stream: Optional[torch.cuda.Stream] = None,
                 on_staged=None):
        """on_staged(slot, *device_tensors): called by put() right after the copies of a batch were enqueued, with
        the copy stream current -- index work that depends on the batch alone (TTEmbeddingBag.prepare) then runs
        beside the compute stream's current step; get() makes the compute stream wait for it."""
        self.on_staged = on_staged
        if depth < 2:
            raise ValueError("HostBatchPipeline: depth must be at least 2 to overlap anything")
        self.device = torch.device(device)
        self.depth = depth
        # the stream the consumers run on: looked up once (torch.cuda.current_stream costs ~10 us)
        self.compute_stream = stream if stream is not None else torch.cuda.current_stream(self.device)
        self.copy_stream = torch.cuda.Stream(self.device)
        self._pinned_ok = set()
        self.slots: List[Optional[List[torch.Tensor]]] = [None] * depth
        self.ready = [torch.cuda.Event() for _ in range(depth)]
        self.freed: List[Optional[torch.cuda.Event]] = [None] * depth
        self.staged = deque()     # slots put() filled, oldest first
        self.in_use = deque()     # slots get() handed out and release() has not seen yet
        self.next_slot = 0
        self.h2d_bytes = 0

    def _buffers(self, slot: int, host: Sequence[torch.Tensor]) -> List[torch.Tensor]:
        bufs = self.slots[slot]
        if bufs is None or len(bufs) != len(host) or any(
                b.shape != h.shape or b.dtype != h.dtype for b, h in zip(bufs, host)):
            bufs = [torch.empty(h.shape, dtype=h.dtype, device=self.device) for h in host]
            self.slots[slot] = bufs
        return bufs

    def put(self, *host: torch.Tensor) -> None:
        if len(self.staged) + len(self.in_use) >= self.depth:
            raise RuntimeError("HostBatchPipeline: all %d slots are staged or in use" % self.depth)
        for h in host:
            if h.data_ptr() not in self._pinned_ok:      # is_pinned() asks the driver: once per buffer
                if h.is_cuda or not h.is_pinned():
                    raise RuntimeError("HostBatchPipeline.put: inputs must be pinned host tensors")
                if len(self._pinned_ok) < 4096:
                    self._pinned_ok.add(h.data_ptr())
        slot = self.next_slot
        self.next_slot = (slot + 1) % self.depth
        bufs = self._buffers(slot, host)
        if self.freed[slot] is not None:
            self.copy_stream.wait_event(self.freed[slot])
        prev = torch.cuda.current_stream(self.device)    # whatever the caller had current, not necessarily ours
        torch.cuda.set_stream(self.copy_stream)          # copy_ takes the current stream
        try:
            for b, h in zip(bufs, host):
                b.copy_(h, non_blocking=True)
                self.h2d_bytes += h.numel() * h.element_size()
            if self.on_staged is not None:
                self.on_staged(slot, *bufs)
        finally:
            torch.cuda.set_stream(prev)
        self.ready[slot].record(self.copy_stream)
        self.staged.append(slot)

    def get(self) -> Tuple[torch.Tensor, ...]:
        if not self.staged:
            raise RuntimeError("HostBatchPipeline.get: nothing staged")
        slot = self.staged.popleft()
        self.compute_stream.wait_event(self.ready[slot])
        self.in_use.append(slot)
        return tuple(self.slots[slot])

    def release(self) -> None:
        slot = self.in_use.popleft()
        ev = self.freed[slot]
        if ev is None:
            ev = self.freed[slot] = torch.cuda.Event()
        ev.record(self.compute_stream)


class DeferredScalars:
    """Device scalars read back `delay` steps late: push() enqueues the device->host copy into
    pinned memory and returns the values whose copies were enqueued `delay` pushes ago (waiting on
    their event only); drain() returns what is still in flight."""

    def __init__(self, device: torch.device, delay: int = 1, dtype=torch.float32,
                 stream: Optional[torch.cuda.Stream] = None):
        self.device = torch.device(device)
        self.delay = delay
        self.stream = stream if stream is not None else torch.cuda.current_stream(self.device)
        self.host = [torch.empty(1, dtype=dtype).pin_memory() for _ in range(delay + 1)]
        self.events = [torch.cuda.Event() for _ in range(delay + 1)]
        self.pending = deque()
        self.slot = 0
        self.d2h_bytes = 0

    def _read(self) -> float:
        s = self.pending.popleft()
        self.events[s].synchronize()
        return float(self.host[s][0])

    def push(self, value: torch.Tensor) -> List[float]:
        out = []
        s = self.slot                              # never a pending one: delay + 1 slots
        self.slot = (s + 1) % (self.delay + 1)
        cur = torch.cuda.current_stream(self.device)     # the stream the copy is issued on
        self.host[s].copy_(value.detach().reshape(1), non_blocking=True)
        self.events[s].record(cur)
        self.d2h_bytes += self.host[s].element_size()
        self.pending.append(s)
        while len(self.pending) > self.delay:
            out.append(self._read())
        return out

    def drain(self) -> List[float]:
        out = []
        while self.pending:
            out.append(self._read())
        return out


class GraphedStep:
    """One training (or inference) step captured into a CUDA graph and replayed: every call on the
    hot path is stream-ordered through the C ABI (plan, group table, forward, backward, fused
    update -- INTEGRATION.md "Streams"), so `fn` may be the whole TTEmbeddingBag step including
    loss.backward().  The inputs are the tensors `fn` closes over; refill them in place (for
    example the slots of a HostBatchPipeline) and call the object: one launch instead of ~25.

    `fn` runs `warmup` times for real before the capture (on a side stream, as torch requires),
    then once more under capture; what it returns are static tensors that every replay
    overwrites.

    Not capturable: a TTEmbeddingBag with use_cache=True after cache_populate() (its
    preprocess_indices_sync reads nnz_tt back with a stream synchronisation and sizes later launches
    with it; tt_embeddings.preprocess_indices_sync raises under capture).  Every replay rewrites the
    shared index-plan workspace, so the host-side plan key is cleared: an eager backward after a replay
    rebuilds its plan instead of trusting the one an earlier eager forward left there."""

    def __init__(self, fn, device: torch.device, warmup: int = 2):
        device = torch.device(device)
        self.device = device
        cur = torch.cuda.current_stream(device)
        side = torch.cuda.Stream(device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(warmup):
                fn()
        cur.wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.outputs = fn()

    def __call__(self):
        import _ttg
        self.graph.replay()
        _ttg.workspace.set_plan(self.device, None)
        return self.outputs
