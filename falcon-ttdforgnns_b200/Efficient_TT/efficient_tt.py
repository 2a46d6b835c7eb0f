"""Eff_TTEmbedding with the reference's surface (Efficient_TT/efficient_tt.py:214-307).

The reference JIT-compiles its extension from a hard-coded /home/... path at import time
(:8-11) and is therefore not importable as shipped; this module binds the same five ops from
`efficient_tt_table` (C ABI of libttg_b200.so) instead.  Forward reuses the core0 x core1
product across rows that share idx // p2 (prefix reuse), backward is the fused SGD update of
the cores (Fused_Extra_Eff_TT_backward) -- no gradients flow to autograd, as in the reference.
"""
import math
from typing import List, Optional, Tuple

import numpy as np
import torch
from torch import nn

import efficient_tt_table as Eff_TT_embedding_cuda
from FBTT.tt_embeddings_ops import suggested_tt_shapes  # same helper as the reference's copy

__all__ = ["Eff_TTEmbedding", "TT_core_function", "suggested_tt_shapes"]


class TT_core_function(torch.autograd.Function):
    """Argument order of Efficient_TT/efficient_tt.py:77-93."""

    @staticmethod
    def forward(ctx, batch_size, table_length, feature_dim, indices, tt_p_shapes, tt_q_shapes,
                tt_ranks, tensor_p_shape, tensor_q_shape, tensor_tt_ranks, sorted_idx, sorted_key,
                learning_rate_or_core, *tt_cores):
        # the reference passes the cores right after sorted_key; an optional float in that slot
        # carries the module's learning rate (the reference hard-codes 0.1 at :140,159)
        if isinstance(learning_rate_or_core, torch.Tensor):
            tt_cores = (learning_rate_or_core,) + tuple(tt_cores)
            lr = 0.1
        else:
            lr = float(learning_rate_or_core)
        ctx.cfg = (batch_size, table_length, feature_dim, list(tt_p_shapes), list(tt_q_shapes),
                   list(tt_ranks), tensor_p_shape, tensor_q_shape, tensor_tt_ranks, lr)
        ctx.tt_cores = tt_cores
        ctx.sorted = (sorted_idx, sorted_key)
        ctx.n_inputs = 13 + len(tt_cores) - (1 if isinstance(learning_rate_or_core, torch.Tensor)
                                             else 0)
        ctx.save_for_backward(indices)
        return Eff_TT_embedding_cuda.Eff_TT_forward(
            batch_size, table_length, feature_dim, indices, tt_p_shapes, tt_q_shapes, tt_ranks,
            tensor_p_shape, tensor_q_shape, tensor_tt_ranks, list(tt_cores))

    @staticmethod
    def backward(ctx, grad_output: torch.Tensor) -> Tuple[Optional[torch.Tensor], ...]:
        (indices,) = ctx.saved_tensors
        (batch_size, table_length, feature_dim, p, q, ranks, tp, tq, tr, lr) = ctx.cfg
        sorted_idx, sorted_key = ctx.sorted
        Eff_TT_embedding_cuda.Fused_Extra_Eff_TT_backward(
            batch_size, table_length, feature_dim, lr, indices, p, q, ranks, tp, tq, tr,
            grad_output, list(ctx.tt_cores), sorted_idx, sorted_key)
        return (None,) * ctx.n_inputs


class Eff_TTEmbedding(nn.Module):
    def __init__(self, num_embeddings: int, embedding_dim: int, tt_ranks: List[int],
                 tt_p_shapes: Optional[List[int]] = None, tt_q_shapes: Optional[List[int]] = None,
                 optimizer: str = "SGD", learning_rate: float = 0.1, weight_dist: str = "uniform",
                 device=0, batch_size=4096) -> None:
        super().__init__()
        self.num_embeddings = num_embeddings
        self.embedding_dim = embedding_dim
        self.num_tt_core = len(tt_ranks) + 1
        self.tt_ranks = [1] + list(tt_ranks) + [1]
        self.batch_size = batch_size
        self.tt_p_shapes = (list(tt_p_shapes) if tt_p_shapes is not None
                            else suggested_tt_shapes(num_embeddings, self.num_tt_core))
        self.tt_q_shapes = (list(tt_q_shapes) if tt_q_shapes is not None
                            else suggested_tt_shapes(embedding_dim, self.num_tt_core))
        assert self.num_tt_core == 3, "Efficient_TT kernels are written for 3 cores"
        assert int(np.prod(self.tt_p_shapes, dtype=np.int64)) >= num_embeddings
        assert int(np.prod(self.tt_q_shapes, dtype=np.int64)) == embedding_dim
        self.optimizer = optimizer
        self.learning_rate = learning_rate
        self.weight_dist = weight_dist
        self.device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        Eff_TT_embedding_cuda.init_cuda(self.device.index or 0, self.tt_q_shapes, self.tt_ranks,
                                        batch_size, embedding_dim)
        self.tt_cores = nn.ParameterList()
        for t in range(self.num_tt_core):
            cols = self.tt_ranks[t] * self.tt_q_shapes[t] * self.tt_ranks[t + 1]
            self.tt_cores.append(nn.Parameter(torch.empty((self.tt_p_shapes[t], cols),
                                                          device=self.device, dtype=torch.float32)))
        self.reset_parameters()
        self.tensor_p_shape = torch.tensor(self.tt_p_shapes, device=self.device)
        self.tensor_q_shape = torch.tensor(self.tt_q_shapes, device=self.device)
        self.tensor_tt_ranks = torch.tensor(self.tt_ranks, device=self.device)

    def reset_parameters(self):
        """`uniform` initialiser of the reference (:277-286); other names leave the cores as
        allocated there, here they fall back to the same initialiser."""
        d = self.num_tt_core
        stddev = math.sqrt(2.0 / (self.num_embeddings + self.embedding_dim))
        rank_term = float(np.prod(np.array(self.tt_ranks, dtype=np.float64) ** (-1.0 / (2 * d))))
        hi = stddev ** (1.0 / d) * rank_term
        with torch.no_grad():
            for c in self.tt_cores:
                c.uniform_(0.0, hi)

    def forward(self, indices, offsets=None, unique=None, inverse=None):
        batch_size = indices.shape[0]
        out = TT_core_function.apply(batch_size, self.num_embeddings, self.embedding_dim, indices,
                                     self.tt_p_shapes, self.tt_q_shapes, self.tt_ranks,
                                     self.tensor_p_shape, self.tensor_q_shape, self.tensor_tt_ranks,
                                     unique, inverse, float(self.learning_rate), *self.tt_cores)
        return out.to(self.device)
