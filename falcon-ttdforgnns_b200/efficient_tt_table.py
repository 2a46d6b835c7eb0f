"""Drop-in for the Efficient_TT extension (JIT name `efficient_tt_table`,
Efficient_TT/efficient_tt.py:8-11; setup.py name `effi_tt_embeddings`).

Same five functions as Efficient_TT/efficient_kernel_wrap.cpp:83-89 with the same argument
orders (:8-80).  Differences in mechanism only: no process-global cudaMalloc'd scratch
(Efficient_TT/efficient_tt_cuda.cu:43-73) -- the scratch is a per-device torch buffer, so several
tables / devices / streams can coexist; integer index math (identical to the reference's float
math wherever that is exact); everything runs on torch's current stream.
"""
import ctypes as C
from typing import List

import torch

import _ttg

_ws = _ttg._Workspace()


def init_cuda(device_id: int, tt_q_shape: List[int], tt_ranks: List[int], batch_size: int,
              feature_dim: int) -> None:
    """Efficient_TT/efficient_tt_cuda.cu:51-73: selects the device; no global state is needed."""
    if not torch.cuda.is_available():
        raise RuntimeError("efficient_tt_table.init_cuda: no CUDA device (there is no CPU path)")
    dev = torch.device(device_id) if not isinstance(device_id, int) else torch.device("cuda",
                                                                                      device_id)
    torch.cuda.set_device(dev)
    _ttg.lib()


def _shape3(tt_p_shapes, tt_q_shapes, tt_ranks):
    if len(tt_p_shapes) != 3:
        raise RuntimeError("Efficient_TT supports exactly 3 cores")
    return _ttg.make_shape(tt_p_shapes, tt_q_shapes, tt_ranks, 1)


def _cores(tt_cores):
    out = []
    for i, c in enumerate(tt_cores):
        c = c.data if isinstance(c, torch.nn.Parameter) else c
        out.append(_ttg.require_cuda(c, "tt_cores[%d]" % i, torch.float32))
    return out


def Eff_TT_forward(batch_size: int, table_length: int, feature_dim: int, index: torch.Tensor,
                   tt_p_shapes: List[int], tt_q_shapes: List[int], tt_ranks: List[int],
                   tensor_p_shape: torch.Tensor, tensor_q_shape: torch.Tensor,
                   tensor_tt_ranks: torch.Tensor, tt_cores: List[torch.Tensor]) -> torch.Tensor:
    """Efficient_TT_forward_cuda (Efficient_TT/efficient_tt_cuda.cu:243-377): out[n] = row(index[n])."""
    cores = _cores(tt_cores)
    dev = cores[0].device
    shape = _shape3(tt_p_shapes, tt_q_shapes, tt_ranks)
    with _ttg.on_device(dev):
        raw = index
        index = _ttg.require_cuda(index.long().contiguous(), "index", torch.int64)
        if index.numel() < batch_size:
            raise RuntimeError("Eff_TT_forward: batch_size exceeds len(index)")
        out = torch.empty((batch_size, feature_dim), dtype=torch.float32, device=dev)
        lib = _ttg.lib()
        ws = _ws.get(dev, lib.ttg_eff_workspace_bytes(C.byref(shape), batch_size))
        cp = _ttg.ptr_array(cores)
        rc = lib.ttg_eff_forward(C.byref(shape), batch_size, _ttg.ptr(index), cp, _ttg.ptr(out),
                                 _ttg.ptr(ws), ws.numel(), _ttg.stream_of(dev))
        _ttg.check(rc, "Eff_TT_forward")
        if index.data_ptr() == raw.data_ptr():  # only trust a plan built from the caller's tensor
            _ws.set_plan(dev, ("eff", index.data_ptr(), index._version, int(batch_size),
                               tuple((c.data_ptr(), c._version) for c in cores)), keep=(index,))
        else:
            _ws.set_plan(dev, None)
    return out


def Fused_Extra_Eff_TT_backward(batch_size: int, table_length: int, feature_dim: int,
                                learning_rate: float, indices: torch.Tensor,
                                tt_p_shapes: List[int], tt_q_shapes: List[int],
                                tt_ranks: List[int], tensor_p_shape: torch.Tensor,
                                tensor_q_shape: torch.Tensor, tensor_tt_ranks: torch.Tensor,
                                d_output: torch.Tensor, tt_cores: List[torch.Tensor],
                                sorted_idx: torch.Tensor = None,
                                sorted_key: torch.Tensor = None) -> None:
    """Fused_Extra_Efficient_TT_backward_sgd_cuda (Efficient_TT/efficient_tt_cuda.cu:1011-1247):
    cores -= lr * gradient, in place.  The reference first sums d_output over duplicate indices
    (unique / inverse, :970-987, efficient_tt.py:132-133); the gradient is linear in d_output, so
    processing every occurrence gives the same result and sorted_idx / sorted_key are not
    needed (they are accepted for signature compatibility)."""
    cores = _cores(tt_cores)
    dev = cores[0].device
    shape = _shape3(tt_p_shapes, tt_q_shapes, tt_ranks)
    with _ttg.on_device(dev):
        indices = _ttg.require_cuda(indices.long().contiguous(), "indices", torch.int64)
        g = d_output.to(torch.float32).contiguous()
        if g.dim() != 2 or g.size(0) < batch_size or g.size(1) != feature_dim:
            raise RuntimeError("Eff_TT_backward: d_output must be [batch_size, feature_dim]")
        lib = _ttg.lib()
        ws = _ws.get(dev, lib.ttg_eff_workspace_bytes(C.byref(shape), batch_size))
        flags = 0
        if _ws.plan(dev) == ("eff", indices.data_ptr(), indices._version, int(batch_size),
                             tuple((c.data_ptr(), c._version) for c in cores)):
            flags = _ttg.FLAG_PLAN_VALID
        cp = _ttg.ptr_array(cores)
        rc = lib.ttg_eff_backward_sgd(C.byref(shape), batch_size, float(learning_rate),
                                      _ttg.ptr(indices), _ttg.ptr(g), cp, _ttg.ptr(ws), ws.numel(),
                                      flags, _ttg.stream_of(dev))
        _ttg.check(rc, "Eff_TT_backward")
        _ws.set_plan(dev, None)   # cores updated in place: plan's group table is stale


def Eff_TT_backward(batch_size, table_length, feature_dim, learning_rate, indices, tt_p_shapes,
                    tt_q_shapes, tt_ranks, tensor_p_shape, tensor_q_shape, tensor_tt_ranks,
                    d_output, tt_cores) -> None:
    """Efficient_TT_backward_sgd_cuda (:496-715).  Dead in the reference (hard-coded 0.1 step,
    :454,472); kept as an alias of the live fused update with the caller's learning rate."""
    Fused_Extra_Eff_TT_backward(batch_size, table_length, feature_dim, learning_rate, indices,
                                tt_p_shapes, tt_q_shapes, tt_ranks, tensor_p_shape, tensor_q_shape,
                                tensor_tt_ranks, d_output, tt_cores)


def Fused_Eff_TT_backward(batch_size, table_length, feature_dim, learning_rate, indices,
                          tt_p_shapes, tt_q_shapes, tt_ranks, tensor_p_shape, tensor_q_shape,
                          tensor_tt_ranks, d_output, tt_cores) -> None:
    """Fused_Efficient_TT_backward_sgd_cuda (:718-904), same update as the live variant."""
    Fused_Extra_Eff_TT_backward(batch_size, table_length, feature_dim, learning_rate, indices,
                                tt_p_shapes, tt_q_shapes, tt_ranks, tensor_p_shape, tensor_q_shape,
                                tensor_tt_ranks, d_output, tt_cores)
