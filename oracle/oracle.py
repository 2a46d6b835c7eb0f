"""numpy front-end of the CPU oracle (oracle/tt_oracle.c) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product (falcon-ttdforgnns_b200/) never does.

Every function takes / returns numpy arrays and mirrors one reference function; the C side
carries the reference file:line citations.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libtt_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "tt_oracle.c")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= os.path.getmtime(src)):
        return _LIB_PATH
    subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), check=True,
                   capture_output=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_hash.restype = C.c_uint32
        _lib.orc_hash.argtypes = [C.c_int64, C.c_int32]
        _lib.orc_preprocess_indices.restype = C.c_int64
        _lib.orc_num_threads.restype = C.c_int
    return _lib


def _p(a, ty):
    return a.ctypes.data_as(C.POINTER(ty))


def _ia(v):
    return (C.c_int * len(v))(*[int(x) for x in v])


def _full_ranks(T, ranks):
    ranks = [int(x) for x in ranks]
    if len(ranks) == T - 1:
        ranks = [1] + ranks + [1]
    assert len(ranks) == T + 1
    return ranks


def _core_ptrs(cores):
    cs = [np.ascontiguousarray(c, dtype=np.float32) for c in cores]
    arr = (C.POINTER(C.c_float) * len(cs))(*[_p(c, C.c_float) for c in cs])
    return cs, arr


def num_threads():
    return int(lib().orc_num_threads())


def use_all_host_threads():
    """OpenMP threads = the cores this process may run on, whatever OMP_NUM_THREADS says."""
    import os
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    lib().orc_set_num_threads(int(n))
    return num_threads()


def tt_forward(p, q, ranks, cores, indices, rowidx, B, tableidx=None, num_tables=1):
    """Reference op tt_forward: zeros[num_tables, B, D] + bag-sum of reconstructed rows."""
    T = len(p)
    r = _full_ranks(T, ranks)
    indices = np.ascontiguousarray(indices, dtype=np.int64)
    rowidx = np.ascontiguousarray(rowidx, dtype=np.int64)
    tableidx = (np.zeros_like(indices) if tableidx is None
                else np.ascontiguousarray(tableidx, dtype=np.int64))
    cs, cp = _core_ptrs(cores)
    D = int(np.prod(q))
    out = np.zeros((num_tables, B, D), dtype=np.float32)
    rc = lib().orc_tt_forward(T, num_tables, _ia(p), _ia(q), _ia(r), C.c_int64(B),
                              C.c_int64(indices.size), _p(indices, C.c_int64),
                              _p(rowidx, C.c_int64), _p(tableidx, C.c_int64), cp,
                              _p(out, C.c_float))
    assert rc == 0
    return out


def tt_forward_f32_rows(p, q, ranks, cores, indices):
    """fp32, one index per row: the CPU baseline bench.py times."""
    T = len(p)
    r = _full_ranks(T, ranks)
    indices = np.ascontiguousarray(indices, dtype=np.int64)
    cs, cp = _core_ptrs(cores)
    out = np.empty((indices.size, int(np.prod(q))), dtype=np.float32)
    rc = lib().orc_tt_forward_f32_rows(T, _ia(p), _ia(q), _ia(r), C.c_int64(indices.size),
                                       _p(indices, C.c_int64), cp, _p(out, C.c_float))
    assert rc == 0
    return out


def tt_backward_f32_rows(p, q, ranks, cores, indices, d_output):
    """fp32, one index per row: the backward half of the CPU baseline bench.py times."""
    T = len(p)
    r = _full_ranks(T, ranks)
    indices = np.ascontiguousarray(indices, dtype=np.int64)
    d_output = np.ascontiguousarray(d_output, dtype=np.float32)
    cs, cp = _core_ptrs(cores)
    outs = [np.zeros_like(c) for c in cs]
    op = (C.POINTER(C.c_float) * T)(*[_p(o, C.c_float) for o in outs])
    rc = lib().orc_tt_backward_f32_rows(T, _ia(p), _ia(q), _ia(r), C.c_int64(indices.size),
                                        _p(indices, C.c_int64), _p(d_output, C.c_float), cp, op)
    assert rc == 0
    return outs


def tt_backward_dense(p, q, ranks, cores, indices, rowidx, d_output, tableidx=None, num_tables=1):
    """Reference op tt_dense_backward: list of d_tt_cores (zeros_like + scatter-add)."""
    T = len(p)
    r = _full_ranks(T, ranks)
    indices = np.ascontiguousarray(indices, dtype=np.int64)
    rowidx = np.ascontiguousarray(rowidx, dtype=np.int64)
    tableidx = (np.zeros_like(indices) if tableidx is None
                else np.ascontiguousarray(tableidx, dtype=np.int64))
    d_output = np.ascontiguousarray(d_output, dtype=np.float32)
    B = d_output.shape[-2]
    cs, cp = _core_ptrs(cores)
    outs = [np.zeros_like(c) for c in cs]
    op = (C.POINTER(C.c_float) * T)(*[_p(o, C.c_float) for o in outs])
    rc = lib().orc_tt_backward_dense(T, num_tables, _ia(p), _ia(q), _ia(r), C.c_int64(B),
                                     C.c_int64(indices.size), _p(indices, C.c_int64),
                                     _p(rowidx, C.c_int64), _p(tableidx, C.c_int64),
                                     _p(d_output, C.c_float), cp, op)
    assert rc == 0
    return outs


def apply_optimizer(p, cols, optim, lr, eps, cores, state, d_cores, rows_limit=None,
                    num_tables=1):
    """In place. optim: 'sgd' | 'adagrad'. rows_limit reproduces the reference launch bug."""
    T = len(p)
    cp = (C.POINTER(C.c_float) * T)(*[_p(c, C.c_float) for c in cores])
    dp = (C.POINTER(C.c_float) * T)(*[_p(c, C.c_float) for c in d_cores])
    sp = None
    if state is not None:
        sp = (C.POINTER(C.c_float) * T)(*[_p(c, C.c_float) for c in state])
    rl = None
    if rows_limit is not None:
        rl = (C.c_int64 * T)(*[int(x) for x in rows_limit])
    lib().orc_apply_optimizer(T, num_tables, _ia(p), _ia(cols), 0 if optim == "sgd" else 1,
                              C.c_float(lr), C.c_float(eps), rl, cp, sp, dp)


def reference_sgd_rows_updated(p, cols):
    """Rows the reference's fused-update launch actually touches:
    grid = ceil(cols/ty) blocks of ty rows, tx = min(1024, p), ty = 1024 // tx
    (FBTT/tt_embeddings_cuda.cu:634-650)."""
    lim = []
    for pt, ct in zip(p, cols):
        tx = min(1024, pt)
        ty = 1024 // tx
        lim.append(min(pt, -(-ct // ty) * ty))
    return lim


def hash32(key, size):
    return int(lib().orc_hash(C.c_int64(int(key)), C.c_int32(int(size))))


def update_cache_state(indices, hashtbl, cache_freq):
    indices = np.ascontiguousarray(indices, dtype=np.int64)
    assert hashtbl.dtype == np.int64 and cache_freq.dtype == np.int64
    lib().orc_update_cache_state(C.c_int64(indices.size), _p(indices, C.c_int64),
                                 C.c_int32(hashtbl.size), _p(hashtbl, C.c_int64),
                                 _p(cache_freq, C.c_int64))


def cache_populate_index(cache_size, hashtbl, cache_freq, cache_state):
    """In place on the three buffers; returns the sorted key list (rows to prefetch are
    sorted_keys[:cache_size])."""
    assert cache_state.dtype == np.int32
    sorted_keys = np.empty_like(hashtbl)
    lib().orc_cache_populate_index(C.c_int32(hashtbl.size), C.c_int32(cache_size),
                                   _p(hashtbl, C.c_int64), _p(cache_freq, C.c_int64),
                                   _p(cache_state, C.c_int32), _p(sorted_keys, C.c_int64))
    return sorted_keys


def preprocess_indices(colidx, offsets, num_tables, warmup, hashtbl, cache_state):
    colidx = np.ascontiguousarray(colidx, dtype=np.int64)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    nnz = colidx.size
    rowidx = np.zeros(nnz, dtype=np.int64)
    tableidx = np.zeros(nnz, dtype=np.int64)
    pc = np.zeros(nnz, dtype=np.int64)
    pr = np.zeros(nnz, dtype=np.int64)
    pl = np.full(nnz, -1, dtype=np.int32)
    ht = hashtbl if hashtbl is not None else np.zeros(1, dtype=np.int64)
    cs = cache_state if cache_state is not None else np.zeros(1, dtype=np.int32)
    n_tt = lib().orc_preprocess_indices(
        C.c_int64(nnz), C.c_int64(offsets.size), _p(colidx, C.c_int64), _p(offsets, C.c_int64),
        C.c_int(num_tables), C.c_int(1 if warmup else 0), C.c_int32(ht.size), _p(ht, C.c_int64),
        _p(cs, C.c_int32), _p(rowidx, C.c_int64), _p(tableidx, C.c_int64), _p(pc, C.c_int64),
        _p(pr, C.c_int64), _p(pl, C.c_int32))
    if warmup or num_tables != 1:
        return colidx, rowidx, tableidx, int(n_tt), None
    return pc, pr, tableidx, int(n_tt), pl


def cache_forward(loc, rowidx, weight, output):
    lib().orc_cache_forward(C.c_int64(loc.size), C.c_int(weight.shape[1]), _p(loc, C.c_int32),
                            _p(rowidx, C.c_int64), _p(weight, C.c_float), _p(output, C.c_float))


def cache_backward(grad_output, loc, rowidx, lr, mode, dst):
    lib().orc_cache_backward(C.c_int64(loc.size), C.c_int(dst.shape[1]),
                             _p(grad_output, C.c_float), _p(loc, C.c_int32),
                             _p(rowidx, C.c_int64), C.c_float(lr), C.c_int(mode),
                             _p(dst, C.c_float))


def cache_backward_rowwise_adagrad(grad_output, loc, rowidx, lr, eps, state, weight):
    lib().orc_cache_backward_rowwise_adagrad(
        C.c_int64(loc.size), C.c_int(weight.shape[1]), _p(grad_output, C.c_float),
        _p(loc, C.c_int32), _p(rowidx, C.c_int64), C.c_float(lr), C.c_float(eps),
        _p(state, C.c_float), _p(weight, C.c_float))


def eff_split(indices, p, use_float):
    indices = np.ascontiguousarray(indices, dtype=np.int64)
    pp = np.ascontiguousarray(p, dtype=np.int64)
    out = np.zeros((indices.size, 4), dtype=np.int32)
    lib().orc_eff_split(C.c_int64(indices.size), _p(indices, C.c_int64), _p(pp, C.c_int64),
                        C.c_int(1 if use_float else 0), _p(out, C.c_int32))
    return out


def spmm_csr_fwd(indptr, indices, x, mean, edge_weight=None):
    indptr = np.ascontiguousarray(indptr, dtype=np.int64)
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.zeros((indptr.size - 1, x.shape[1]), dtype=np.float32)
    ew = None if edge_weight is None else _p(np.ascontiguousarray(edge_weight, np.float32),
                                             C.c_float)
    lib().orc_spmm_csr_fwd(C.c_int64(indptr.size - 1), C.c_int(x.shape[1]), _p(indptr, C.c_int64),
                           _p(indices, C.c_int32), ew, C.c_int(1 if mean else 0),
                           _p(x, C.c_float), _p(out, C.c_float))
    return out


def spmm_csr_bwd(indptr, indices, dout, num_src, mean, edge_weight=None):
    indptr = np.ascontiguousarray(indptr, dtype=np.int64)
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    dout = np.ascontiguousarray(dout, dtype=np.float32)
    dx = np.zeros((num_src, dout.shape[1]), dtype=np.float32)
    ew = None if edge_weight is None else _p(np.ascontiguousarray(edge_weight, np.float32),
                                             C.c_float)
    lib().orc_spmm_csr_bwd(C.c_int64(indptr.size - 1), C.c_int64(num_src), C.c_int(dout.shape[1]),
                           _p(indptr, C.c_int64), _p(indices, C.c_int32), ew,
                           C.c_int(1 if mean else 0), _p(dout, C.c_float), _p(dx, C.c_float))
    return dx
