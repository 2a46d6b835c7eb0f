"""Drop-in for the reference's compiled extension module `tt_embeddings`.

Same 11 function names, argument orders and return types as the pybind module registered at
FBTT/tt_embeddings.cpp:131-161 (signatures :13-129), so `import tt_embeddings` inside
FBTT/tt_embeddings_ops.py (and anything else that calls the raw ops) keeps working when this
directory is first on sys.path.  Each function validates its tensors, allocates outputs with
torch and forwards raw device pointers to the C ABI (include/ttg_b200.h) on torch's current
stream.  Errors surface as RuntimeError, like TORCH_CHECK in the reference.

`batch_count` is accepted and ignored: in the reference it only sizes the scratch of the chunk
loop (FBTT/tt_embeddings_cuda.cu:1015-1027); here one launch covers the whole batch.
"""
import ctypes as C
from typing import List, Optional, Tuple

import torch

import _ttg

__all__ = [
    "tt_forward", "tt_dense_backward", "tt_sgd_backward", "tt_adagrad_backward",
    "update_cache_state", "cache_populate", "preprocess_indices_sync", "cache_forward",
    "cache_backward_sgd", "cache_backward_dense", "cache_backward_rowwise_adagrad_approx",
]

_scratch = _ttg._Workspace()  # everything that is not the sorted TT plan
_grad_scratch = {}             # (device, table shape) -> d_cores of the fused-update ops

# test / debugging knob: OR-ed into the flags of tt_forward / tt_*_backward
# (_ttg.FLAG_FORCE_GENERIC selects the shape-generic kernels)
EXTRA_FLAGS = 0


def _cores_ok(tt_cores, T):
    if len(tt_cores) != T:
        raise RuntimeError("tt_embeddings: expected %d tt_cores, got %d" % (T, len(tt_cores)))
    cores = []
    for i, c in enumerate(tt_cores):
        c = c.data if isinstance(c, torch.nn.Parameter) else c
        cores.append(_ttg.require_cuda(c, "tt_cores[%d]" % i, torch.float32))
    return cores


def _plan_tag():
    # a plan (and group table) built under one set of flags is not reused under another
    return "tt-%d" % EXTRA_FLAGS


def _shape_tuple(p, q, r, num_tables):
    return (tuple(int(x) for x in p), tuple(int(x) for x in q), tuple(int(x) for x in r),
            int(num_tables))


def tt_plan(num_tables: int, B: int, tt_p_shapes: List[int], tt_q_shapes: List[int], tt_ranks: List[int],
            nnz: int, indices: torch.Tensor, rowidx: torch.Tensor, tableidx: torch.Tensor, slot: int,
            device=None) -> bool:
    """Build the index plan of a batch ahead of its tt_forward, on torch's CURRENT stream, into plan slot `slot`
    (0 / 1; the batch being processed meanwhile uses the other one).  Not an op of the reference: its forward
    splits the indices inside every call (FBTT/tt_embeddings_cuda.cu:757-921); here a data loader that knows
    the next batch prepares it beside the host-to-device copy (pipeline.HostBatchPipeline(on_staged=...)).
    tt_forward / tt_*_backward recognise the prepared batch by its tensors and skip the plan.  The caller orders
    the streams (the forward's stream must wait for this one).  Returns False when the shape has no index plan
    (shape-generic kernels): nothing was prepared, the forward simply plans itself."""
    dev = indices.device if device is None else device
    shape = _ttg.make_shape(tt_p_shapes, tt_q_shapes, tt_ranks, num_tables)
    nnz = int(nnz)
    if nnz == 0:
        return False
    _ttg.require_cuda(indices, "indices", torch.int64)
    _ttg.require_cuda(rowidx, "rowidx", torch.int64)
    _ttg.require_cuda(tableidx, "tableidx", torch.int64)
    with _ttg.on_device(dev):
        lib = _ttg.lib()
        nbytes = _ttg.tt_workspace_bytes(shape, B, nnz)
        ws = _ttg.workspace.get(dev, nbytes)
        # the plan is built BESIDE the batch in flight: only under the layout that batch uses (same shape, B, nnz)
        if not _ttg.workspace.same_layout(dev, (shape.key, int(B), nnz), adopt=False):
            return False
        _ttg.workspace.same_layout(dev, (shape.key, int(B), nnz))
        flags = EXTRA_FLAGS | (_ttg.FLAG_PLAN_SLOT1 if slot else 0)
        rc = lib.ttg_tt_plan(C.byref(shape), B, nnz, _ttg.ptr(indices), _ttg.ptr(rowidx), _ttg.ptr(tableidx),
                             _ttg.ptr(ws), ws.numel(), flags, _ttg.stream_of(dev))
        if rc == -4:                                # TTG_ENOTSUP
            return False
        _ttg.check(rc, "tt_plan")
        _ttg.workspace.set_ready(dev, slot, _ttg.index_key_of(_plan_tag(), indices, rowidx, nnz, B, shape.key),
                                 keep=(indices, rowidx))
        # whatever batch's full plan (plan + table) this slot held is gone
        cur = _ttg.workspace.plan(dev)
        if cur is not None and cur[-1] == slot:
            _ttg.workspace.set_plan(dev, None)
    return True


def tt_forward(batch_count: int, num_tables: int, B: int, D: int, tt_p_shapes: List[int],
               tt_q_shapes: List[int], tt_ranks: List[int], L: torch.Tensor, nnz: int,
               indices: torch.Tensor, rowidx: torch.Tensor, tableidx: torch.Tensor,
               tt_cores: List[torch.Tensor]) -> torch.Tensor:
    """tt_embeddings_forward_cuda (FBTT/tt_embeddings_cuda.cu:967-1081)."""
    cores = _cores_ok(tt_cores, len(tt_p_shapes))
    dev = cores[0].device
    shape = _ttg.make_shape(tt_p_shapes, tt_q_shapes, tt_ranks, num_tables)
    if cores[0].size(0) != num_tables:
        raise RuntimeError("tt_forward: num_tables does not match tt_cores[0].size(0)")
    with _ttg.on_device(dev):
        output = torch.empty((num_tables, B, D), dtype=torch.float32, device=dev)
        nnz = int(nnz)
        if nnz > 0:
            _ttg.require_cuda(indices, "indices", torch.int64)
            _ttg.require_cuda(rowidx, "rowidx", torch.int64)
            _ttg.require_cuda(tableidx, "tableidx", torch.int64)
            if min(indices.numel(), rowidx.numel(), tableidx.numel()) < nnz:
                raise RuntimeError("tt_forward: nnz exceeds the index arrays")
        lib = _ttg.lib()
        nbytes = _ttg.tt_workspace_bytes(shape, B, nnz)
        ws = _ttg.workspace.get(dev, nbytes)
        cp = _ttg.ptr_array(cores)
        flags, slot = EXTRA_FLAGS, 0
        _ttg.workspace.same_layout(dev, (shape.key, int(B), nnz))
        if nnz > 0:
            ready = _ttg.workspace.ready_slot(dev, _ttg.index_key_of(_plan_tag(), indices, rowidx, nnz, B,
                                                                     shape.key))
            if ready is not None:                  # tt_plan prepared this batch
                slot = ready
                flags |= _ttg.FLAG_PLAN_READY | (_ttg.FLAG_PLAN_SLOT1 if slot else 0)
            else:                                  # planning in slot 0 overwrites what was prepared there
                _ttg.workspace.drop_ready(dev, 0)
        rc = lib.ttg_tt_forward(C.byref(shape), B, nnz, _ttg.ptr(indices), _ttg.ptr(rowidx),
                                _ttg.ptr(tableidx), cp, _ttg.ptr(output), _ttg.ptr(ws),
                                ws.numel(), flags, _ttg.stream_of(dev))
        _ttg.check(rc, "tt_forward")
        if nnz > 0:
            _ttg.workspace.set_plan(dev, _ttg.plan_key_of(
                _plan_tag(), indices, rowidx, nnz, B,
                shape.key, tt_cores) + (slot,),   # the caller's tensors: a Parameter's own version counter
                keep=(indices, rowidx))
    return output


def tt_rows_range(first_row: int, num_rows: int, tt_p_shapes: List[int], tt_q_shapes: List[int],
                  tt_ranks: List[int], tt_cores: List[torch.Tensor]) -> Optional[torch.Tensor]:
    """Rows [first_row, first_row + num_rows) of a single-table TT matrix, in order, without index
    arrays (ttg_tt_rows_range; not an op of the reference, which gets the same rows from
    tt_forward(arange) -- gcn_gat_partition.py:93-96).  Returns None when the shape has no
    tensor-core kernels; the caller then uses tt_forward on an explicit range."""
    cores = _cores_ok(tt_cores, len(tt_p_shapes))
    if len(tt_p_shapes) != 3 or cores[0].size(0) != 1:
        return None
    dev = cores[0].device
    shape = _ttg.make_shape(tt_p_shapes, tt_q_shapes, tt_ranks, 1)
    D = 1
    for x in tt_q_shapes:
        D *= int(x)
    with _ttg.on_device(dev):
        lib = _ttg.lib()
        nbytes = lib.ttg_tt_rows_range_workspace_bytes(C.byref(shape))
        if nbytes == 0:
            return None
        out = torch.empty((int(num_rows), D), dtype=torch.float32, device=dev)
        ws = _ttg.workspace.get(dev, nbytes)
        _ttg.workspace.same_layout(dev, ("rows_range", shape.key))   # its own carving: prepared plans are gone
        _ttg.workspace.set_plan(dev, None)          # the workspace no longer holds a batch's plan
        rc = lib.ttg_tt_rows_range(C.byref(shape), int(first_row), int(num_rows), _ttg.ptr_array(cores),
                                   _ttg.ptr(out), _ttg.ptr(ws), ws.numel(), EXTRA_FLAGS,
                                   _ttg.stream_of(dev))
        if rc == -4:                                # TTG_ENOTSUP
            return None
        _ttg.check(rc, "tt_rows_range")
    return out


def _backward(optim, D, lr, eps, tt_p_shapes, tt_q_shapes, tt_ranks, nnz, indices, rowidx,
              tableidx, d_output, optimizer_state, tt_cores):
    cores = _cores_ok(tt_cores, len(tt_p_shapes))
    dev = cores[0].device
    num_tables = cores[0].size(0)
    shape = _ttg.make_shape(tt_p_shapes, tt_q_shapes, tt_ranks, num_tables)
    with _ttg.on_device(dev):
        if not isinstance(d_output, torch.Tensor) or not d_output.is_cuda:
            raise RuntimeError("tt_backward: d_output must be a CUDA tensor")
        d_output = d_output.to(torch.float32).contiguous()
        if d_output.dim() != 3 or d_output.size(0) != num_tables or d_output.size(2) != D:
            raise RuntimeError("tt_backward: d_output must be [num_tables, B, D]")
        B = d_output.size(1)
        nnz = int(nnz)
        if optim == _ttg.OPTIM_DENSE:
            # one allocation, views per core: a data-parallel caller all-reduces the flat buffer
            # without a concatenation (dp.flatten recognises the layout); 16-byte aligned views
            sizes = [(c.numel() + 3) // 4 * 4 for c in cores]
            flat = torch.empty(sum(sizes), dtype=torch.float32, device=dev)
            d_cores, off = [], 0
            for c, n in zip(cores, sizes):
                d_cores.append(flat[off:off + c.numel()].view_as(c))
                off += n
        else:
            # fused update: the dense gradients are scratch, keep one set per (device, table)
            ck = (dev.index, shape.key)
            d_cores = _grad_scratch.get(ck)
            if d_cores is None or any(g.shape != c.shape for g, c in zip(d_cores, cores)):
                d_cores = [torch.empty_like(c) for c in cores]
                if len(_grad_scratch) < 64:
                    _grad_scratch[ck] = d_cores
        if nnz > 0:
            _ttg.require_cuda(indices, "indices", torch.int64)
            _ttg.require_cuda(rowidx, "rowidx", torch.int64)
            _ttg.require_cuda(tableidx, "tableidx", torch.int64)
        states = None
        sp = None
        if optim == _ttg.OPTIM_ADAGRAD:
            states = [_ttg.require_cuda(s, "optimizer_state", torch.float32)
                      for s in optimizer_state]
            for s, c in zip(states, cores):
                if s.shape != c.shape:
                    raise RuntimeError("tt_adagrad_backward: optimizer_state shape mismatch")
            sp = _ttg.ptr_array(states)
        lib = _ttg.lib()
        nbytes = _ttg.tt_workspace_bytes(shape, B, nnz)
        ws = _ttg.workspace.get(dev, nbytes)
        _ttg.workspace.same_layout(dev, (shape.key, int(B), nnz))
        flags = 0
        key = None
        if nnz > 0:
            key = _ttg.plan_key_of(_plan_tag(), indices, rowidx, nnz, B,
                                   shape.key,
                                   tt_cores)
            cur = _ttg.workspace.plan(dev)
            if cur is not None and cur[:-1] == key:
                flags |= _ttg.FLAG_PLAN_VALID  # the forward's sort is still in the workspace
                if cur[-1]:
                    flags |= _ttg.FLAG_PLAN_SLOT1
            else:
                _ttg.workspace.drop_ready(dev, 0)  # this call plans in slot 0
        cp = _ttg.ptr_array(cores)
        dp = _ttg.ptr_array(d_cores)
        rc = lib.ttg_tt_backward(C.byref(shape), optim, float(lr), float(eps), B, nnz,
                                 _ttg.ptr(indices), _ttg.ptr(rowidx), _ttg.ptr(tableidx),
                                 _ttg.ptr(d_output), cp, sp, dp, _ttg.ptr(ws), ws.numel(),
                                 flags | EXTRA_FLAGS, _ttg.stream_of(dev))
        _ttg.check(rc, "tt_backward")
        if nnz > 0:
            if optim != _ttg.OPTIM_DENSE:
                # the cores were updated in place by the kernel: the group table is stale
                _ttg.workspace.set_plan(dev, None)
            elif not flags:
                _ttg.workspace.set_plan(dev, key + (0,), keep=(indices, rowidx))
    return d_cores


def tt_dense_backward(batch_count: int, D: int, tt_p_shapes: List[int], tt_q_shapes: List[int],
                      tt_ranks: List[int], L: torch.Tensor, nnz: int, indices: torch.Tensor,
                      rowidx: torch.Tensor, tableidx: torch.Tensor, d_output: torch.Tensor,
                      tt_cores: List[torch.Tensor]) -> List[torch.Tensor]:
    """tt_embeddings_backward_dense_cuda (FBTT/tt_embeddings_cuda.cu:656-686)."""
    return _backward(_ttg.OPTIM_DENSE, D, 0.0, 0.0, tt_p_shapes, tt_q_shapes, tt_ranks, nnz, indices,
                     rowidx, tableidx, d_output, None, tt_cores)


def tt_sgd_backward(batch_count: int, D: int, learning_rate: float, tt_p_shapes: List[int],
                    tt_q_shapes: List[int], tt_ranks: List[int], L: torch.Tensor, nnz: int,
                    indices: torch.Tensor, rowidx: torch.Tensor, tableidx: torch.Tensor,
                    d_output: torch.Tensor, tt_cores: List[torch.Tensor]) -> None:
    """tt_embeddings_backward_sgd_cuda (FBTT/tt_embeddings_cuda.cu:688-719): in place."""
    _backward(_ttg.OPTIM_SGD, D, learning_rate, 0.0, tt_p_shapes, tt_q_shapes, tt_ranks, nnz,
              indices, rowidx, tableidx, d_output, None, tt_cores)


def tt_adagrad_backward(batch_count: int, D: int, learning_rate: float, eps: float,
                        tt_p_shapes: List[int], tt_q_shapes: List[int], tt_ranks: List[int],
                        L: torch.Tensor, nnz: int, indices: torch.Tensor, rowidx: torch.Tensor,
                        tableidx: torch.Tensor, d_output: torch.Tensor,
                        optimizer_state: List[torch.Tensor],
                        tt_cores: List[torch.Tensor]) -> None:
    """tt_embeddings_backward_adagrad_cuda (FBTT/tt_embeddings_cuda.cu:721-754): in place."""
    _backward(_ttg.OPTIM_ADAGRAD, D, learning_rate, eps, tt_p_shapes, tt_q_shapes, tt_ranks, nnz,
              indices, rowidx, tableidx, d_output, optimizer_state, tt_cores)


def update_cache_state(indices: torch.Tensor, hashtbl: torch.Tensor,
                       cache_freq: torch.Tensor) -> None:
    """update_cache_state_cuda (FBTT/tt_embeddings_cuda.cu:1097-1119)."""
    if indices.numel() == 0:
        return
    _ttg.require_cuda(indices, "indices", torch.int64)
    _ttg.require_cuda(hashtbl, "hashtbl", torch.int64)
    _ttg.require_cuda(cache_freq, "cache_freq", torch.int64)
    if hashtbl.numel() == 0 or hashtbl.numel() != cache_freq.numel():
        raise RuntimeError("update_cache_state: hashtbl and cache_freq must be non-empty and of "
                           "equal length")
    dev = indices.device
    with _ttg.on_device(dev):
        rc = _ttg.lib().ttg_update_cache_state(indices.numel(), _ttg.ptr(indices), hashtbl.numel(),
                                               _ttg.ptr(hashtbl), _ttg.ptr(cache_freq),
                                               _ttg.stream_of(dev))
        _ttg.check(rc, "update_cache_state")


def cache_populate(num_embeddings: int, tt_p_shapes: List[int], tt_q_shapes: List[int],
                   tt_ranks: List[int], tt_cores: List[torch.Tensor], L: torch.Tensor,
                   hashtbl: torch.Tensor, cache_freq: torch.Tensor, cache_state: torch.Tensor,
                   cache_weight: torch.Tensor) -> None:
    """cache_populate_cuda (FBTT/tt_embeddings_cuda.cu:1270-1347)."""
    cores = _cores_ok(list(tt_cores), len(tt_p_shapes))
    _ttg.require_cuda(hashtbl, "hashtbl", torch.int64)
    _ttg.require_cuda(cache_freq, "cache_freq", torch.int64)
    _ttg.require_cuda(cache_state, "cache_state", torch.int32)
    cw = cache_weight.data if isinstance(cache_weight, torch.nn.Parameter) else cache_weight
    _ttg.require_cuda(cw, "cache_weight", torch.float32)
    if hashtbl.numel() == 0 or hashtbl.numel() != cache_freq.numel():
        raise RuntimeError("cache_populate: hashtbl/cache_freq size mismatch")
    if hashtbl.numel() < cw.size(0):
        raise RuntimeError("cache_populate: hashtbl smaller than the cache")
    dev = hashtbl.device
    shape = _ttg.make_shape(tt_p_shapes, tt_q_shapes, tt_ranks, cores[0].size(0))
    with _ttg.on_device(dev):
        lib = _ttg.lib()
        nbytes = lib.ttg_cache_populate_workspace_bytes(C.byref(shape), hashtbl.numel(), cw.size(0))
        ws = _scratch.get(dev, nbytes)
        cp = _ttg.ptr_array(cores)
        rc = lib.ttg_cache_populate(C.byref(shape), cp, hashtbl.numel(), _ttg.ptr(hashtbl),
                                    _ttg.ptr(cache_freq), _ttg.ptr(cache_state), cw.size(0),
                                    _ttg.ptr(cw), _ttg.ptr(ws), ws.numel(), _ttg.stream_of(dev))
        _ttg.check(rc, "cache_populate")


def preprocess_indices_sync(colidx: torch.Tensor, offsets: torch.Tensor, num_tables: int,
                            warmup: bool, hashtbl: torch.Tensor, cache_state: torch.Tensor
                            ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, int,
                                       Optional[torch.Tensor]]:
    """preprocess_indices_sync_cuda (FBTT/tt_embeddings_cuda.cu:1388-1507)."""
    _ttg.require_cuda(colidx, "colidx", torch.int64)
    _ttg.require_cuda(offsets, "offsets", torch.int64)
    dev = colidx.device
    nnz = colidx.numel()
    with _ttg.on_device(dev):
        rowidx = torch.empty_like(colidx)
        tableidx = torch.empty_like(colidx)
        if nnz == 0:
            return colidx, rowidx, tableidx, 0, None
        lookup = (not warmup) and num_tables == 1
        if lookup and torch.cuda.is_current_stream_capturing():
            raise RuntimeError("preprocess_indices_sync: the cached / uncached split returns a host integer "
                               "(stream synchronisation) and cannot be captured in a CUDA graph; capture with "
                               "use_cache=False or before cache_populate()")
        part_col = part_row = part_loc = None
        if lookup:
            _ttg.require_cuda(hashtbl, "hashtbl", torch.int64)
            _ttg.require_cuda(cache_state, "cache_state", torch.int32)
            part_col = torch.empty_like(colidx)
            part_row = torch.empty_like(colidx)
            part_loc = torch.empty(nnz, dtype=torch.int32, device=dev)
        lib = _ttg.lib()
        ws = _scratch.get(dev, lib.ttg_preprocess_workspace_bytes(nnz))
        n_tt = C.c_int32(0)
        rc = lib.ttg_preprocess_indices(
            nnz, offsets.numel(), _ttg.ptr(colidx), _ttg.ptr(offsets), int(num_tables),
            1 if warmup else 0, hashtbl.numel() if lookup else 0,
            _ttg.ptr(hashtbl) if lookup else None, _ttg.ptr(cache_state) if lookup else None,
            _ttg.ptr(rowidx), _ttg.ptr(tableidx), _ttg.ptr(part_col), _ttg.ptr(part_row),
            _ttg.ptr(part_loc), C.byref(n_tt), _ttg.ptr(ws), ws.numel(), _ttg.stream_of(dev))
        _ttg.check(rc, "preprocess_indices_sync")
    if not lookup:
        return colidx, rowidx, tableidx, nnz, None
    return part_col, part_row, tableidx, int(n_tt.value), part_loc


def cache_mark(colidx: torch.Tensor, hashtbl: torch.Tensor, cache_state: torch.Tensor
               ) -> Tuple[torch.Tensor, torch.Tensor]:
    """(tt_colidx, cache_locations) of ttg_cache_mark: the cached / uncached split in place -- cached entries
    become id -1 for the TT ops, uncached ones location -1 for the cache ops -- with no host count, hence no
    stream synchronisation (not an op of the reference: its split is preprocess_indices_sync above)."""
    _ttg.require_cuda(colidx, "colidx", torch.int64)
    _ttg.require_cuda(hashtbl, "hashtbl", torch.int64)
    _ttg.require_cuda(cache_state, "cache_state", torch.int32)
    dev = colidx.device
    with _ttg.on_device(dev):
        tt_col = torch.empty_like(colidx)
        loc = torch.empty(colidx.numel(), dtype=torch.int32, device=dev)
        rc = _ttg.lib().ttg_cache_mark(colidx.numel(), _ttg.ptr(colidx), hashtbl.numel(), _ttg.ptr(hashtbl),
                                       _ttg.ptr(cache_state), _ttg.ptr(tt_col), _ttg.ptr(loc),
                                       _ttg.stream_of(dev))
        _ttg.check(rc, "cache_mark")
    return tt_col, loc


def _cache_args(nnz, grad_or_out, cache_locations, rowidx, cache_weight):
    _ttg.require_cuda(cache_locations, "cache_locations", torch.int32)
    _ttg.require_cuda(rowidx, "rowidx", torch.int64)
    cw = cache_weight.data if isinstance(cache_weight, torch.nn.Parameter) else cache_weight
    _ttg.require_cuda(cw, "cache_weight", torch.float32)
    if min(cache_locations.numel(), rowidx.numel()) < nnz:
        raise RuntimeError("cache op: nnz exceeds cache_locations / rowidx")
    return cw


def cache_forward(B: int, nnz: int, cache_locations: torch.Tensor, rowidx: torch.Tensor,
                  cache_weight: torch.Tensor, output: torch.Tensor) -> None:
    """cache_forward_cuda (FBTT/tt_embeddings_cuda.cu:1551-1583): accumulates into output."""
    if B <= 0:
        raise RuntimeError("cache_forward: B must be positive")
    cw = _cache_args(nnz, output, cache_locations, rowidx, cache_weight)
    _ttg.require_cuda(output, "output", torch.float32)
    dev = output.device
    with _ttg.on_device(dev):
        rc = _ttg.lib().ttg_cache_forward(int(nnz), cw.size(1), _ttg.ptr(cache_locations),
                                          _ttg.ptr(rowidx), _ttg.ptr(cw), _ttg.ptr(output),
                                          _ttg.stream_of(dev))
        _ttg.check(rc, "cache_forward")


def cache_backward_sgd(nnz: int, grad_output: torch.Tensor, cache_locations: torch.Tensor,
                       rowidx: torch.Tensor, learning_rate: float,
                       cache_weight: torch.Tensor) -> None:
    """cache_backward_sgd_cuda (FBTT/tt_embeddings_cuda.cu:1634-1668)."""
    cw = _cache_args(nnz, grad_output, cache_locations, rowidx, cache_weight)
    dev = cw.device
    with _ttg.on_device(dev):
        g = grad_output.to(torch.float32).contiguous()
        rc = _ttg.lib().ttg_cache_backward_sgd(int(nnz), cw.size(1), _ttg.ptr(g),
                                               _ttg.ptr(cache_locations), _ttg.ptr(rowidx),
                                               float(learning_rate), _ttg.ptr(cw),
                                               _ttg.stream_of(dev))
        _ttg.check(rc, "cache_backward_sgd")


def cache_backward_dense(nnz: int, grad_output: torch.Tensor, cache_locations: torch.Tensor,
                         rowidx: torch.Tensor, learning_rate: float,
                         cache_weight: torch.Tensor) -> torch.Tensor:
    """cache_backward_dense_cuda (FBTT/tt_embeddings_cuda.cu:1710-1744)."""
    cw = _cache_args(nnz, grad_output, cache_locations, rowidx, cache_weight)
    dev = cw.device
    with _ttg.on_device(dev):
        grad = torch.zeros_like(cw)
        g = grad_output.to(torch.float32).contiguous()
        rc = _ttg.lib().ttg_cache_backward_dense(int(nnz), cw.size(1), _ttg.ptr(g),
                                                 _ttg.ptr(cache_locations), _ttg.ptr(rowidx),
                                                 _ttg.ptr(grad), _ttg.stream_of(dev))
        _ttg.check(rc, "cache_backward_dense")
    return grad


def cache_backward_rowwise_adagrad_approx(nnz: int, grad_output: torch.Tensor,
                                          cache_locations: torch.Tensor, rowidx: torch.Tensor,
                                          learning_rate: float, eps: float,
                                          cache_optimizer_state: torch.Tensor,
                                          cache_weight: torch.Tensor) -> None:
    """cache_backward_rowwise_adagrad_approx_cuda (FBTT/tt_embeddings_cuda.cu:1808-1846)."""
    cw = _cache_args(nnz, grad_output, cache_locations, rowidx, cache_weight)
    _ttg.require_cuda(cache_optimizer_state, "cache_optimizer_state", torch.float32)
    dev = cw.device
    with _ttg.on_device(dev):
        g = grad_output.to(torch.float32).contiguous()
        rc = _ttg.lib().ttg_cache_backward_rowwise_adagrad_approx(
            int(nnz), cw.size(1), _ttg.ptr(g), _ttg.ptr(cache_locations), _ttg.ptr(rowidx),
            float(learning_rate), float(eps), _ttg.ptr(cache_optimizer_state), _ttg.ptr(cw),
            _ttg.stream_of(dev))
        _ttg.check(rc, "cache_backward_rowwise_adagrad_approx")
