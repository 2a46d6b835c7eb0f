"""Neighbour aggregation that consumes the reconstructed rows (north star part (d)).

The reference's models call DGL 2.1.0 (un-vendored): dglnn.SAGEConv(in, out, 'mean') three
times (gnn_model.py:78-81, forward :206-217) and dglnn.GraphConv(norm='both',
allow_zero_in_degree=True) (gnn_model.py:287).  DGL is not part of this image, so the sparse part
is restated here over a minimal bipartite block:

    Block.indptr  int64 [num_dst + 1]   CSR by destination
    Block.indices int32 [num_edges]     source ids local to the block
    destination nodes are the first num_dst source nodes (h_dst = h[:num_dst], gnn_model.py:211)

and runs through ttg_spmm_csr_fwd / ttg_spmm_csr_bwd of the C ABI.  The dense parts (Linear,
bias) stay in torch: they are library GEMMs, not part of the path.
"""
import ctypes as C
from dataclasses import dataclass
from typing import Optional, Tuple

import torch
from torch import nn

import _ttg


@dataclass
class Block:
    indptr: torch.Tensor       # int64 [num_dst + 1]
    indices: torch.Tensor      # int32 [E]
    num_src: int
    num_dst: int

    def num_dst_nodes(self):
        return self.num_dst

    def num_src_nodes(self):
        return self.num_src

    def in_degrees(self):
        return (self.indptr[1:] - self.indptr[:-1])

    def out_degrees(self):
        return torch.bincount(self.indices.long(), minlength=self.num_src)

    def int(self):
        return self

    def to(self, device):
        return Block(self.indptr.to(device), self.indices.to(device), self.num_src, self.num_dst)


class _SpMM(torch.autograd.Function):
    @staticmethod
    def forward(ctx, indptr, indices, x, num_dst, mean, edge_weight):
        _ttg.require_cuda(indptr, "indptr", torch.int64)
        _ttg.require_cuda(indices, "indices", torch.int32)
        x = _ttg.require_cuda(x.contiguous(), "x", torch.float32)
        dev = x.device
        F = x.size(1)
        with _ttg.on_device(dev):
            out = torch.empty((num_dst, F), dtype=torch.float32, device=dev)
            rc = _ttg.lib().ttg_spmm_csr_fwd(num_dst, F, _ttg.ptr(indptr), _ttg.ptr(indices),
                                             _ttg.ptr(edge_weight), 1 if mean else 0, _ttg.ptr(x),
                                             _ttg.ptr(out), _ttg.stream_of(dev))
            _ttg.check(rc, "spmm_csr_fwd")
        ctx.save_for_backward(indptr, indices, edge_weight)
        ctx.cfg = (x.size(0), num_dst, F, mean)
        return out

    @staticmethod
    def backward(ctx, dout):
        indptr, indices, edge_weight = ctx.saved_tensors
        num_src, num_dst, F, mean = ctx.cfg
        dev = dout.device
        with _ttg.on_device(dev):
            dout = dout.to(torch.float32).contiguous()
            dx = torch.zeros((num_src, F), dtype=torch.float32, device=dev)
            rc = _ttg.lib().ttg_spmm_csr_bwd(num_dst, F, _ttg.ptr(indptr), _ttg.ptr(indices),
                                             _ttg.ptr(edge_weight), 1 if mean else 0,
                                             _ttg.ptr(dout), _ttg.ptr(dx), _ttg.stream_of(dev))
            _ttg.check(rc, "spmm_csr_bwd")
        return None, None, dx, None, None, None


def aggregate(block: Block, x: torch.Tensor, mean: bool,
              edge_weight: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[v] = (1/deg(v) if mean) * sum_{u in N_in(v)} w_uv * x[u]; rows without in-edges are 0."""
    return _SpMM.apply(block.indptr, block.indices, x, block.num_dst, mean, edge_weight)


class SAGEConv(nn.Module):
    """GraphSAGE layer with DGL 2.1 `SAGEConv(in, out, 'mean')` semantics:
    out = fc_self(h_dst) + fc_neigh(mean_{u in N(v)} h_src[u]) + bias, the neighbour Linear being
    applied BEFORE the aggregation iff in_feats > out_feats (so the last 256 -> 47 layer gathers
    47-wide rows and the first 100 -> 256 layer gathers the raw reconstructed rows)."""

    def __init__(self, in_feats: int, out_feats: int, aggregator_type: str = "mean",
                 bias: bool = True):
        super().__init__()
        if aggregator_type != "mean":
            raise NotImplementedError("only the 'mean' aggregator of the reference models")
        self.in_feats, self.out_feats = in_feats, out_feats
        self.fc_self = nn.Linear(in_feats, out_feats, bias=False)
        self.fc_neigh = nn.Linear(in_feats, out_feats, bias=False)
        self.bias = nn.Parameter(torch.zeros(out_feats)) if bias else None
        gain = nn.init.calculate_gain("relu")
        nn.init.xavier_uniform_(self.fc_self.weight, gain=gain)
        nn.init.xavier_uniform_(self.fc_neigh.weight, gain=gain)

    def forward(self, block: Block, feat: Tuple[torch.Tensor, torch.Tensor]) -> torch.Tensor:
        h_src, h_dst = feat if isinstance(feat, tuple) else (feat, feat[:block.num_dst])
        if self.in_feats > self.out_feats:
            h_neigh = aggregate(block, self.fc_neigh(h_src), mean=True)
        else:
            h_neigh = self.fc_neigh(aggregate(block, h_src, mean=True))
        out = self.fc_self(h_dst) + h_neigh
        if self.bias is not None:
            out = out + self.bias
        return out


class GraphConv(nn.Module):
    """GCN layer with DGL `GraphConv(norm='both', allow_zero_in_degree=True)` semantics:
    h = D_in^-1/2 * A * (D_out^-1/2 * x) with degrees clamped to >= 1, weight applied before the
    aggregation iff in_feats > out_feats."""

    def __init__(self, in_feats: int, out_feats: int, norm: str = "both", bias: bool = True):
        super().__init__()
        if norm != "both":
            raise NotImplementedError("only norm='both' of the reference GCN")
        self.in_feats, self.out_feats = in_feats, out_feats
        self.weight = nn.Parameter(torch.empty(in_feats, out_feats))
        nn.init.xavier_uniform_(self.weight)
        self.bias = nn.Parameter(torch.zeros(out_feats)) if bias else None

    def forward(self, block: Block, feat: torch.Tensor) -> torch.Tensor:
        out_deg = block.out_degrees().clamp(min=1).to(torch.float32)
        x = feat * out_deg.pow(-0.5).unsqueeze(1)
        if self.in_feats > self.out_feats:
            rst = aggregate(block, x @ self.weight, mean=False)
        else:
            rst = aggregate(block, x, mean=False) @ self.weight
        in_deg = block.in_degrees().clamp(min=1).to(torch.float32)
        rst = rst * in_deg.pow(-0.5).unsqueeze(1)
        if self.bias is not None:
            rst = rst + self.bias
        return rst
