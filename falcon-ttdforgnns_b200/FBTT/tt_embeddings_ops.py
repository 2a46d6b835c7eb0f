"""B200-native TTEmbeddingBag with the reference's module surface.

Mirrors FBTT/tt_embeddings_ops.py of JoshuaQSH/FALCON-TTDforGNNs: the names OptimType,
BufferList, tt_matrix_to_full, TTLookupFunction, suggested_tt_shapes,
TableBatchedTTEmbeddingBag and TTEmbeddingBag, their constructor / forward arguments
(:446-464, :923-958, :837-903, :960-965), buffers (L, hashtbl, cache_freq, cache_state,
optimizer_state<i>) and parameters (tt_cores, cache_weight), so that gnn_model.py:113-125,
gcn_gat_partition.py:217-227 and sage_profiler.py:210-221 construct and call it unchanged.
All computation goes through the `tt_embeddings` op module next to this package, i.e. through
the C ABI of libttg_b200.so; there is no eager / CPU path.
"""
import enum
import itertools
import logging
import math
from typing import List, Optional, Sequence

import numpy as np
import torch
import tt_embeddings
from torch import nn

__all__ = ["OptimType", "BufferList", "tt_matrix_to_full", "TTLookupFunction",
           "suggested_tt_shapes", "TableBatchedTTEmbeddingBag", "TTEmbeddingBag"]


@enum.unique
class OptimType(enum.Enum):
    """Optimizer selector (reference :18-33). Only SGD / EXACT_SGD (fused core -= lr*g) and the
    Adagrad family have fused kernels; with sparse=False the dense gradients go to autograd."""
    SGD = "sgd"
    EXACT_SGD = "exact_sgd"
    LAMB = "lamb"
    ADAM = "adam"
    EXACT_ADAGRAD = "exact_adagrad"
    EXACT_ROWWISE_ADAGRAD = "exact_row_wise_adagrad"
    LARS_SGD = "lars_sgd"
    PARTIAL_ROWWISE_ADAM = "partial_row_wise_adam"
    PARTIAL_ROWWISE_LAMB = "partial_row_wise_lamb"

    def __str__(self):
        return self.value


_SGD_LIKE = (OptimType.SGD, OptimType.EXACT_SGD)


class BufferList(nn.Module):
    """nn.ParameterList for buffers: entries are registered as `<name><i>` (reference :36-77)."""

    def __init__(self, name: str, buffers: Optional[Sequence[torch.Tensor]] = None):
        super().__init__()
        self._name = name
        self._length = 0
        if buffers is not None:
            self.extend(buffers)

    def append(self, buffer: torch.Tensor) -> "BufferList":
        self.register_buffer("%s%d" % (self._name, self._length), buffer)
        self._length += 1
        return self

    def extend(self, buffers: Sequence[torch.Tensor]) -> "BufferList":
        for b in buffers:
            self.append(b)
        return self

    def __len__(self) -> int:
        return self._length

    def __getitem__(self, index: int) -> torch.Tensor:
        if index < 0:
            index += self._length
        return getattr(self, "%s%d" % (self._name, index))

    def __iter__(self):
        return (self[i] for i in range(self._length))


def tt_matrix_to_full(tt_p_shapes: List[int], tt_q_shapes: List[int], tt_ranks: List[int],
                      tt_cores: Sequence[torch.Tensor],
                      tt_permute: Optional[List[int]] = None) -> torch.Tensor:
    """Dense [prod(p), prod(q)] matrix of a TT table (pure PyTorch; reference :80-127).

    Each core arrives as a tensor whose 4-D view has its axes stored in the order given by
    `tt_permute` relative to (r_t, p_t, q_t, r_{t+1}); the module stores [p_t, r_t, q_t, r_{t+1}]
    (num_tables == 1), i.e. tt_permute = [1, 0, 2, 3].
    """
    d = len(tt_p_shapes)
    ranks = list(tt_ranks)
    if len(ranks) == d - 1:
        ranks = [1] + ranks + [1]
    canon = []
    for t, core in enumerate(tt_cores):
        dims = [ranks[t], tt_p_shapes[t], tt_q_shapes[t], ranks[t + 1]]
        if tt_permute is None:
            c = torch.squeeze(core)
            c = c.reshape(dims)
        else:
            stored = [dims[a] for a in tt_permute]
            c = core.reshape(stored).permute(*tt_permute).contiguous()
        if list(c.shape) != dims:
            raise AssertionError("core %d has shape %s, expected %s" % (t, list(c.shape), dims))
        canon.append(c)
    acc = canon[0].reshape(-1, ranks[1])                     # [(p0 q0), r1]
    for t in range(1, d):
        acc = acc.reshape(-1, ranks[t]) @ canon[t].reshape(ranks[t], -1)
    # acc is [p0, q0, p1, q1, ...]: bring the p's in front of the q's
    inter = list(itertools.chain.from_iterable(zip(tt_p_shapes, tt_q_shapes)))
    order = list(range(0, 2 * d, 2)) + list(range(1, 2 * d, 2))
    full = acc.reshape(inter).permute(*order).contiguous()
    return full.reshape(int(np.prod(tt_p_shapes)), int(np.prod(tt_q_shapes))).float()


class TTLookupFunction(torch.autograd.Function):
    """Autograd node around the fused lookup (argument order of the reference, :133-156)."""

    @staticmethod
    def forward(ctx, B, D, tt_p_shapes, tt_q_shapes, tt_ranks, L, nnz_tt, nnz_cached, indices,
                rowidx, tableidx, optimizer, learning_rate, eps, sparse, cache_locations,
                cache_optimizer_state, cache_weight, optimizer_state, batch_count, *tt_cores):
        # the cached entries follow the nnz_tt TT entries (preprocess_indices_sync), or -- unpartitioned lists of
        # tt_embeddings.cache_mark, nnz_tt == nnz_cached == len(indices) -- both kinds span the whole list
        c0 = 0 if (nnz_cached > 0 and nnz_tt + nnz_cached > indices.numel()) else int(nnz_tt)
        ctx.cfg = (D, list(tt_p_shapes), list(tt_q_shapes), list(tt_ranks), optimizer,
                   learning_rate, eps, sparse, int(nnz_tt), int(nnz_cached), batch_count, c0)
        ctx.tt_cores = tt_cores
        ctx.optimizer_state = optimizer_state
        ctx.save_for_backward(L, indices, rowidx, tableidx, cache_locations,
                              cache_optimizer_state, cache_weight)
        out = tt_embeddings.tt_forward(batch_count, tt_cores[0].size(0), B, D, tt_p_shapes,
                                       tt_q_shapes, tt_ranks, L, nnz_tt, indices, rowidx, tableidx,
                                       list(tt_cores))
        if nnz_cached > 0:
            tt_embeddings.cache_forward(B, nnz_cached, cache_locations[c0:], rowidx[c0:],
                                        cache_weight, out)
        return out

    @staticmethod
    def backward(ctx, d_output):
        (D, p, q, ranks, optimizer, lr, eps, sparse, nnz_tt, nnz_cached, batch_count, c0) = ctx.cfg
        (L, indices, rowidx, tableidx, cache_locations, cache_optimizer_state,
         cache_weight) = ctx.saved_tensors
        cores = list(ctx.tt_cores)
        n_fixed = 20  # non-core arguments of forward()
        grads = [None] * n_fixed
        if sparse:
            # fused update, nothing flows back to autograd (reference :228-312)
            if optimizer in _SGD_LIKE:
                tt_embeddings.tt_sgd_backward(batch_count, D, lr, p, q, ranks, L, nnz_tt, indices,
                                              rowidx, tableidx, d_output, cores)
                if nnz_cached > 0:
                    tt_embeddings.cache_backward_sgd(nnz_cached, d_output,
                                                     cache_locations[c0:], rowidx[c0:], lr,
                                                     cache_weight)
            else:
                tt_embeddings.tt_adagrad_backward(batch_count, D, lr, eps, p, q, ranks, L, nnz_tt,
                                                  indices, rowidx, tableidx, d_output,
                                                  ctx.optimizer_state, cores)
                if nnz_cached > 0:
                    tt_embeddings.cache_backward_rowwise_adagrad_approx(
                        nnz_cached, d_output, cache_locations[c0:], rowidx[c0:], lr, eps,
                        cache_optimizer_state, cache_weight)
            return tuple(grads + [None] * len(cores))
        # dense gradients for an external optimizer (reference :313-366)
        d_cores = tt_embeddings.tt_dense_backward(batch_count, D, p, q, ranks, L, nnz_tt, indices,
                                                  rowidx, tableidx, d_output, cores)
        if nnz_cached > 0:
            grads[17] = tt_embeddings.cache_backward_dense(nnz_cached, d_output,
                                                           cache_locations[c0:],
                                                           rowidx[c0:], lr, cache_weight)
        return tuple(grads + list(d_cores))


# ---------------------------------------------------------------------------------------------
# shape suggestion (reference :369-429) without the sympy / scipy dependency
# ---------------------------------------------------------------------------------------------
def _prime_factors(n: int) -> List[int]:
    out, f = [], 2
    while f * f <= n:
        while n % f == 0:
            out.append(f)
            n //= f
        f += 1 if f == 2 else 2
    if n > 1:
        out.append(n)
    return out


def _factorizations(n: int, d: int, lo: int = 2):
    """Non-decreasing d-tuples of integers >= lo with product n."""
    if d == 1:
        if n >= lo:
            yield (n,)
        return
    f = lo
    while f ** d <= n:
        if n % f == 0:
            for rest in _factorizations(n // f, d - 1, f):
                yield (f,) + rest
        f += 1


def _entropy(xs) -> float:
    tot = float(sum(xs))
    return -sum((x / tot) * math.log(x / tot) for x in xs if x > 0)


def _interleave(xs: Sequence[int]) -> List[int]:
    half = len(xs) // 2
    front, back = list(xs[:half]), list(xs[half:])
    out = []
    for i in range(max(len(front), len(back))):
        if i < len(front):
            out.append(front[i])
        if i < len(back):
            out.append(back[i])
    return out


def _auto_shape(n: int, d: int) -> List[int]:
    primes = _prime_factors(n)
    if len(primes) < d:  # not enough prime factors: pad with ones
        cands = [tuple(sorted(primes + [1] * (d - len(primes))))]
    else:
        cands = list(_factorizations(n, d))
    best = max(cands, key=lambda c: (_entropy(c), c))
    return _interleave(sorted(best))


def suggested_tt_shapes(n: int, d: int = 3, allow_round_up: bool = True) -> List[int]:
    """Most balanced factorisation of n (or of n rounded up to a power of ten multiple) into d
    factors; balance is measured by the entropy of the normalised factors."""
    if not allow_round_up:
        return _auto_shape(n, d)
    best_shape, best_w = None, -1.0
    for k in range(len(str(n))):
        rounded = int(math.ceil(n / 10 ** k)) * 10 ** k
        shape = _auto_shape(rounded, d)
        w = _entropy(shape)
        if w > best_w:
            best_shape, best_w = shape, w
    return best_shape


# ---------------------------------------------------------------------------------------------
# modules
# ---------------------------------------------------------------------------------------------
class TableBatchedTTEmbeddingBag(nn.Module):
    """`num_tables` TT tables of identical shape looked up in one pass (reference :432-915)."""

    __constants__ = ["num_tables", "num_embeddings", "embedding_dim", "tt_shape", "tt_rank"]

    def __init__(self, num_tables: int, num_embeddings: int, embedding_dim: int,
                 tt_ranks: List[int], tt_p_shapes: Optional[List[int]] = None,
                 tt_q_shapes: Optional[List[int]] = None, optimizer: OptimType = OptimType.SGD,
                 learning_rate: float = 0.1, eps: float = 1.0e-10, sparse: bool = True,
                 use_cache: bool = False, cache_size: int = 0, hashtbl_size: int = 0,
                 weight_dist: str = "approx-normal", enforce_embedding_dim: bool = False,
                 batch_count: int = 1000) -> None:
        super().__init__()
        if not torch.cuda.is_available():
            raise AssertionError("TTEmbeddingBag needs a CUDA device: there is no CPU path")
        assert num_tables > 0 and num_embeddings > 0 and embedding_dim > 0
        assert num_tables == 1 or not use_cache, "cannot use cache when num_tables != 1"
        n_cores = len(tt_ranks) + 1
        self.batch_count = batch_count
        self.tt_p_shapes = (list(tt_p_shapes) if tt_p_shapes is not None
                            else suggested_tt_shapes(num_embeddings, n_cores))
        self.tt_q_shapes = (list(tt_q_shapes) if tt_q_shapes is not None
                            else suggested_tt_shapes(embedding_dim, n_cores,
                                                     allow_round_up=not enforce_embedding_dim))
        assert 2 <= len(self.tt_p_shapes) <= 4
        assert len(self.tt_p_shapes) == n_cores == len(self.tt_q_shapes)
        assert all(v > 0 for v in self.tt_p_shapes + self.tt_q_shapes + list(tt_ranks))
        assert int(np.prod(self.tt_p_shapes, dtype=np.int64)) >= num_embeddings
        assert int(np.prod(self.tt_q_shapes, dtype=np.int64)) == embedding_dim
        self.num_tables = num_tables
        self.tt_ndim = n_cores
        self.num_embeddings = num_embeddings
        self.embedding_dim = embedding_dim
        self.tt_ranks = [1] + list(tt_ranks) + [1]
        self.sparse = sparse
        self.optimizer = optimizer
        self.learning_rate = learning_rate
        self.eps = eps
        logging.info("Creating TTEmbeddingBag p=%s q=%s ranks=%s sparse=%s optimizer=%s lr=%s "
                     "eps=%s use_cache=%s cache_size=%s hashtbl_size=%s", self.tt_p_shapes,
                     self.tt_q_shapes, self.tt_ranks, sparse, optimizer, learning_rate, eps,
                     use_cache, cache_size, hashtbl_size)
        dev = torch.device("cuda", torch.cuda.current_device())
        # L[t] = prod(p[t+1:]) : radix weights of the index split
        radix = [int(np.prod(self.tt_p_shapes[t + 1:], dtype=np.int64)) for t in range(n_cores)]
        self.register_buffer("L", torch.tensor(radix, dtype=torch.int64))
        self.tt_cores = nn.ParameterList()
        self.optimizer_state = BufferList("optimizer_state")
        for t in range(n_cores):
            cols = self.tt_ranks[t] * self.tt_q_shapes[t] * self.tt_ranks[t + 1]
            core = torch.empty((num_tables, self.tt_p_shapes[t], cols), device=dev,
                               dtype=torch.float32)
            self.tt_cores.append(nn.Parameter(core))
            state_shape = 0 if optimizer in _SGD_LIKE else tuple(core.shape)
            self.optimizer_state.append(torch.zeros(state_shape, device=dev, dtype=torch.float32))
        self.reset_parameters(weight_dist)
        self.use_cache = use_cache
        if use_cache:
            if cache_size <= 0:
                cache_size = int(0.1 * num_embeddings)
            if hashtbl_size <= 0:
                hashtbl_size = num_embeddings
            assert hashtbl_size >= cache_size
            self.register_buffer("hashtbl", torch.full((hashtbl_size,), -1, device=dev,
                                                       dtype=torch.int64))
            self.register_buffer("cache_freq", torch.zeros(hashtbl_size, device=dev,
                                                           dtype=torch.int64))
            self.register_buffer("cache_state", torch.full((hashtbl_size,), -1, device=dev,
                                                           dtype=torch.int32))
            self.cache_weight = nn.Parameter(torch.zeros((cache_size, embedding_dim), device=dev,
                                                         dtype=torch.float32))
            if sparse and optimizer not in _SGD_LIKE:
                shape = ((cache_size, embedding_dim) if optimizer == OptimType.EXACT_ADAGRAD
                         else (cache_size,))
                # the reference allocates this on the CPU (:598-601), which its own kernel
                # cannot read; we keep it with the cache
                self.register_buffer("cache_optimizer_state",
                                     torch.zeros(shape, device=dev, dtype=torch.float32))
            else:
                self.cache_optimizer_state = None
        else:
            self.register_buffer("hashtbl", torch.empty(0, device=dev, dtype=torch.int64))
            self.register_buffer("cache_state", torch.empty(0, device=dev, dtype=torch.int32))
            self.cache_optimizer_state = None
            self.cache_weight = None
        self.warmup = True
        # the cached / uncached split without a host count (tt_embeddings.cache_mark); False = the reference's
        # data flow (partitioned lists, one stream synchronisation per forward)
        self.split_on_device = True

    # -- weights ------------------------------------------------------------------------------
    def full_weight(self) -> torch.Tensor:
        assert self.num_tables == 1, "full_weight() only supported for num_tables == 1 for now"
        if not torch.is_grad_enabled():   # no autograd wanted: the row kernel on implicit keys
            w = self.rows_range(0, int(np.prod(self.tt_p_shapes)))
            if w is not None:
                return w
        return tt_matrix_to_full(self.tt_p_shapes, self.tt_q_shapes, self.tt_ranks,
                                 list(self.tt_cores), [1, 0, 2, 3])

    def rows_range(self, first_row: int, num_rows: int) -> Optional[torch.Tensor]:
        """Rows [first_row, first_row + num_rows) in order, not differentiable: what the
        full-graph drivers and SAGE.inference request as forward(arange(N), arange(N + 1))
        (gcn_gat_partition.py:93-96, gnn_model.py:228-231), without index arrays or plan.
        None when the shape has no tensor-core kernels (use forward on an arange then)."""
        if self.num_tables != 1 or self.tt_ndim != 3:
            return None
        return tt_embeddings.tt_rows_range(first_row, num_rows, self.tt_p_shapes, self.tt_q_shapes,
                                           self.tt_ranks, [c.data for c in self.tt_cores])

    def reset_parameters(self, weight_dist: str) -> None:
        """Core initialisers (reference :629-808)."""
        kinds = ("uniform", "naive-uniform", "normal", "approx-uniform", "approx-normal")
        assert weight_dist in kinds
        d = self.tt_ndim
        with torch.no_grad():
            if weight_dist == "uniform":
                # Var of the product ~ Glorot variance 2/(N+D), spread over d cores and ranks
                stddev = math.sqrt(2.0 / (self.num_embeddings + self.embedding_dim))
                rank_term = float(np.prod(np.array(self.tt_ranks, dtype=np.float64)
                                          ** (-1.0 / (2 * d))))
                hi = stddev ** (1.0 / d) * rank_term
                for c in self.tt_cores:
                    c.uniform_(0.0, hi)
            elif weight_dist == "naive-uniform":
                for c in self.tt_cores:
                    c.uniform_(0.0, 1.0 / math.sqrt(self.num_embeddings))
            elif weight_dist == "normal":
                sigma = 1.0 / math.sqrt(self.num_embeddings)
                for c in self.tt_cores:
                    c.normal_(0.0, sigma)
                    c.mul_(1.0 / self.tt_ranks[0])
            elif weight_dist == "approx-normal":
                # |x| >= 2 tails of N(0,1), scaled so the d-fold product is ~ N(0, 1/(3N));
                # rejection sampling, vectorised (the reference loops per element :663-675)
                scale = (1.0 / math.sqrt(3 * self.num_embeddings)) ** (1.0 / 3.0)
                for c in self.tt_cores:
                    w = np.random.normal(size=tuple(c.shape)).astype(np.float32)
                    bad = np.abs(w) < 2
                    while bad.any():
                        w[bad] = np.random.normal(size=int(bad.sum())).astype(np.float32)
                        bad = np.abs(w) < 2
                    c.copy_(torch.from_numpy(w * np.float32(scale)))
            else:
                self._init_approx_uniform()

    def _init_approx_uniform(self, sigma: float = 0.01, grid: int = 15, width: float = 0.7 / 30):
        """"Flat saw-tooth" initialiser of the reference (:676-808) for 3 cores, one table: head
        core ~ N(1/sqrt(r1), sigma), middle ~ N(1/sqrt(r1), sigma) with one saw-tooth entry per
        (p, q) slot on a random even r2 column, tail ~ N(0, sigma) with one saw-tooth entry per
        slot on a random odd r2 row; everything scaled by N^(-1/6)."""
        assert self.tt_ndim == 3 and self.num_tables == 1
        rng = np.random
        r = self.tt_ranks
        scale = 1.0 / (math.sqrt(self.num_embeddings) ** (1.0 / 3.0))

        def saw(n):
            j = rng.randint(-(grid - 1), grid, n)
            return j * (1.0 / grid) + (-width / 2.0 + width * rng.rand(n))

        p, q = self.tt_p_shapes, self.tt_q_shapes
        head = 1.0 / math.sqrt(r[1]) + rng.randn(r[0], p[0], q[0], r[1]) * sigma
        mid_scale = 1.0 / math.sqrt(r[1])
        mid = (mid_scale + rng.randn(r[1], p[1] * q[1], r[2]) * sigma)
        vals = saw(p[1] * q[1]) / mid_scale
        cols = rng.randint(0, (r[2] + 1) // 2, p[1] * q[1]) * 2
        rows = rng.randint(0, r[1], p[1] * q[1])
        slots = np.arange(p[1] * q[1])
        mid[:, slots, cols] = rng.randn(r[1], p[1] * q[1]) * (sigma * sigma / mid_scale)
        mid[rows, slots, cols] = vals
        mid = mid.reshape(r[1], p[1], q[1], r[2])
        tail = (rng.randn(r[2], p[2] * q[2]) * sigma)
        odd = rng.randint(0, r[2] // 2, p[2] * q[2]) * 2 + 1 if r[2] > 1 else np.zeros(p[2] * q[2], int)
        tail[odd, np.arange(p[2] * q[2])] = saw(p[2] * q[2])
        tail = tail.reshape(r[2], p[2], q[2], r[3])
        for t, w in enumerate((head, mid, tail)):
            w = (w * scale).astype(np.float32).transpose(1, 0, 2, 3).reshape(1, p[t], -1)
            self.tt_cores[t].copy_(torch.from_numpy(np.ascontiguousarray(w)))

    # -- cache --------------------------------------------------------------------------------
    def reset_cache(self):
        if self.use_cache:
            self.hashtbl.fill_(-1)
            self.cache_freq.fill_(0)
            self.cache_state.fill_(-1)
            self.warmup = True

    def cache_populate(self):
        """Freeze the LFU statistics into a cache of reconstructed rows (reference :816-830)."""
        if self.use_cache:
            tt_embeddings.cache_populate(self.num_embeddings, self.tt_p_shapes, self.tt_q_shapes,
                                         self.tt_ranks, list(self.tt_cores), self.L, self.hashtbl,
                                         self.cache_freq, self.cache_state, self.cache_weight)
            self.warmup = False

    def update_cache(self, indices: torch.Tensor):
        if self.use_cache:
            tt_embeddings.update_cache_state(indices, self.hashtbl, self.cache_freq)

    # -- lookup -------------------------------------------------------------------------------
    def prepare(self, indices: torch.Tensor, offsets: torch.Tensor, slot: int = 0) -> bool:
        """Index work of a LATER forward(indices, offsets), on torch's current stream: the bag -> row map and the
        sorted index plan (tt_embeddings.tt_plan) go into plan slot `slot` while the batch in flight uses the
        other one.  Not part of the reference's module (DGL's DataLoader prefetches its batches,
        sage_dgl_partition.py:141-154; the index split stays inside its forward); a loader that stages batches
        ahead calls this beside the host-to-device copy (pipeline.HostBatchPipeline(on_staged=...)).  forward()
        recognises the prepared tensors (same storage, unmodified) and skips both steps.  Only without the LFU
        split (use_cache False or still warming up); returns False when nothing was prepared."""
        if self.use_cache and not self.warmup:
            return False
        if indices.dtype != torch.int64 or offsets.dtype != torch.int64 or not indices.is_contiguous() \
                or not offsets.is_contiguous():
            return False
        (_, rowidx, tableidx, n_tt, _) = tt_embeddings.preprocess_indices_sync(
            indices, offsets, self.num_tables, True, self.hashtbl, self.cache_state)
        B = (offsets.numel() - 1) // self.num_tables
        ok = tt_embeddings.tt_plan(self.num_tables, B, self.tt_p_shapes, self.tt_q_shapes, self.tt_ranks,
                                   n_tt, indices, rowidx, tableidx, slot)
        if not ok:
            return False
        if not hasattr(self, "_prepared"):
            self._prepared = {}
        self._prepared[slot] = ((indices.data_ptr(), indices._version, indices.numel(), offsets.data_ptr(),
                                 offsets._version, offsets.numel()), indices, offsets, rowidx, tableidx)
        return True

    def _prepared_for(self, indices, offsets):
        for slot, (key, i, o, rowidx, tableidx) in getattr(self, "_prepared", {}).items():
            if key == (indices.data_ptr(), indices._version, indices.numel(), offsets.data_ptr(),
                       offsets._version, offsets.numel()):
                return rowidx, tableidx
        return None

    def forward(self, indices: torch.Tensor, offsets: torch.Tensor,
                warmup: bool = True) -> torch.Tensor:
        """[num_tables, B, D] bag sums.  `warmup` is accepted for signature compatibility; like
        the reference (:862) the module's own self.warmup decides whether the cache is used."""
        indices = indices.long().contiguous()
        offsets = offsets.long().contiguous()
        self.update_cache(indices)
        cached = self.use_cache and not self.warmup
        prepared = None if cached else self._prepared_for(indices, offsets)
        if prepared is not None:
            rowidx, tableidx = prepared
            n_tt, cache_locations = indices.numel(), None
            n_cached = 0
        elif cached and self.num_tables == 1 and self.split_on_device:
            # the split stays on the device (tt_embeddings.cache_mark): cached entries become id -1 for the TT
            # kernels, uncached ones location -1 for the cache kernels, both over the whole list -- no host count,
            # no stream synchronisation, and the step can be captured in a CUDA graph with the cache on
            (_, rowidx, tableidx, _, _) = tt_embeddings.preprocess_indices_sync(
                indices, offsets, self.num_tables, True, self.hashtbl, self.cache_state)
            indices, cache_locations = tt_embeddings.cache_mark(indices, self.hashtbl, self.cache_state)
            n_tt = n_cached = indices.numel()
        else:
            (indices, rowidx, tableidx, n_tt, cache_locations) = tt_embeddings.preprocess_indices_sync(
                indices, offsets, self.num_tables, self.warmup, self.hashtbl, self.cache_state)
            n_cached = indices.numel() - n_tt
        return TTLookupFunction.apply(
            (offsets.numel() - 1) // self.num_tables, self.embedding_dim, self.tt_p_shapes,
            self.tt_q_shapes, self.tt_ranks, self.L, n_tt, n_cached, indices, rowidx, tableidx,
            self.optimizer, self.learning_rate, self.eps, self.sparse, cache_locations,
            self.cache_optimizer_state, self.cache_weight, list(self.optimizer_state),
            self.batch_count, *self.tt_cores)

    def set_learning_rate(self, lr: float) -> None:
        self.learning_rate = lr

    def get_params(self) -> List[torch.Tensor]:
        params = list(self.tt_cores)
        if self.use_cache:
            params.append(self.cache_weight)
        return params


class TTEmbeddingBag(TableBatchedTTEmbeddingBag):
    """Exactly one TT table; forward returns [B, D] (reference :918-965)."""

    def __init__(self, num_embeddings: int, embedding_dim: int, tt_ranks: List[int],
                 tt_p_shapes: Optional[List[int]] = None, tt_q_shapes: Optional[List[int]] = None,
                 optimizer: OptimType = OptimType.SGD, learning_rate: float = 0.1,
                 eps: float = 1.0e-10, sparse: bool = True, use_cache: bool = True,
                 cache_size: int = 0, hashtbl_size: int = 0, weight_dist: str = "approx-normal",
                 enforce_embedding_dim: bool = False, batch_count: int = 1000) -> None:
        super().__init__(1, num_embeddings, embedding_dim, tt_ranks, tt_p_shapes, tt_q_shapes,
                         optimizer, learning_rate, eps, sparse, use_cache, cache_size, hashtbl_size,
                         weight_dist, enforce_embedding_dim, batch_count)

    def forward(self, indices: torch.Tensor, offsets: torch.Tensor,
                warmup: bool = True) -> torch.Tensor:
        # squeeze, not [0]: autograd's select-backward allocates a zero-filled [1, B, D] tensor and
        # copies the gradient into it (three passes over B * D floats per step); the adjoint of
        # squeeze is a view
        return super().forward(indices, offsets, warmup).squeeze(0)
