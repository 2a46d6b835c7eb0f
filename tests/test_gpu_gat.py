"""GAT sparse kernels: edge softmax and attention-weighted per-head aggregation
through the C ABI, forward and backward, against a float64 torch restatement of the DGL
semantics the reference relies on (edge_softmax over in-edges; u_mul_e + sum) -- DGL itself is
un-vendored, so this row is parity-unpinned against the third-party code."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _block(rng, num_src, num_dst, max_deg):
    import gnn_ops
    deg = rng.integers(0, max_deg + 1, size=num_dst)
    deg[0] = 0                      # a destination without in-edges
    deg[1] = 70                     # more edges than a warp has lanes
    indptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int64)
    indices = rng.integers(0, num_src, size=int(indptr[-1])).astype(np.int32)
    blk = gnn_ops.Block(torch.from_numpy(indptr).to(DEV), torch.from_numpy(indices).to(DEV),
                        num_src, num_dst)
    dst_of_edge = np.repeat(np.arange(num_dst), deg)
    return blk, torch.from_numpy(dst_of_edge), torch.from_numpy(indices.astype(np.int64))


def _ref_softmax(score, dst_of_edge, num_dst):
    # float64, per (dst, head) segments
    m = torch.full((num_dst, score.size(1)), -float("inf"), dtype=score.dtype)
    m = m.scatter_reduce(0, dst_of_edge[:, None].expand_as(score), score, "amax")
    ex = (score - m[dst_of_edge]).exp()
    z = torch.zeros(num_dst, score.size(1), dtype=score.dtype).index_add_(0, dst_of_edge, ex)
    return ex / z[dst_of_edge]


@pytest.mark.parametrize("gather", [True, False], ids=["gather_backward", "atomic_backward"])
@pytest.mark.parametrize("H,F", [(1, 64), (3, 256), (4, 47), (8, 5)])
def test_edge_softmax_and_aggregation_forward_backward(ttg_lib, H, F, gather, monkeypatch):
    """(both backward passes of the aggregation: the gather over the transposed block and the atomics)"""
    import gnn_ops
    monkeypatch.setattr(gnn_ops, "GATHER_BACKWARD", gather)
    monkeypatch.setattr(gnn_ops, "_gather_backward", lambda block: gather)     # also on this non-square block
    rng = np.random.default_rng(H * 100 + F)
    num_src, num_dst = 900, 400
    blk, dst_e, src_e = _block(rng, num_src, num_dst, 12)
    E = src_e.numel()
    g = torch.Generator().manual_seed(1)
    score64 = (torch.randn(E, H, generator=g, dtype=torch.float64) * 2).requires_grad_(True)
    ft64 = torch.randn(num_src, H, F, generator=g, dtype=torch.float64).requires_grad_(True)
    a64 = _ref_softmax(score64, dst_e, num_dst)
    out64 = torch.zeros(num_dst, H, F, dtype=torch.float64).index_add_(
        0, dst_e, a64[:, :, None] * ft64[src_e])
    w = torch.randn(num_dst, H, F, generator=g, dtype=torch.float64)
    (out64 * w).sum().backward()

    score = score64.detach().float().to(DEV).requires_grad_(True)
    ft = ft64.detach().float().to(DEV).requires_grad_(True)
    a = gnn_ops.edge_softmax(blk, score)
    out = gnn_ops.attention_aggregate(blk, a, ft)
    assert float((a.detach().cpu().double() - a64.detach()).abs().max()) < 1e-6
    assert float((out.detach().cpu().double() - out64.detach()).abs().max() / out64.abs().max()) < 1e-5
    assert float(out[0].abs().max()) == 0.0                       # no in-edges -> zero row
    (out * w.float().to(DEV)).sum().backward()
    for got, want in ((score.grad, score64.grad), (ft.grad, ft64.grad)):
        assert float((got.cpu().double() - want).abs().max() / want.abs().max()) < 2e-5


def test_gatconv_matches_dense_restatement(ttg_lib):
    """GATConv on a block (tuple features, norm='both', residual) against the same formulas in
    float64 torch with scatter ops; gradients of every parameter."""
    import gnn_ops
    rng = np.random.default_rng(5)
    num_src, num_dst, fin, H, F = 500, 200, 48, 3, 16
    blk, dst_e, src_e = _block(rng, num_src, num_dst, 9)
    torch.manual_seed(0)
    conv = gnn_ops.GATConv((fin, fin), F, num_heads=H, residual=True, norm="both").to(DEV)
    x_src = torch.randn(num_src, fin, device=DEV)
    x_dst = x_src[:num_dst]
    out = conv(blk, (x_src, x_dst))
    assert out.shape == (num_dst, H, F)
    # float64 restatement
    P = {k: v.detach().cpu().double().requires_grad_(True) for k, v in conv.named_parameters()}
    xs, xd = x_src.cpu().double(), x_dst.cpu().double()
    fs = (xs @ P["fc_src.weight"].T).view(-1, H, F)
    fd = (xd @ P["fc_dst.weight"].T).view(-1, H, F)
    outdeg = torch.bincount(src_e, minlength=num_src).double().clamp(min=1)
    fs = fs * outdeg.pow(-0.5).view(-1, 1, 1)
    el, er = (fs * P["attn_l"]).sum(-1), (fd * P["attn_r"]).sum(-1)
    e = torch.nn.functional.leaky_relu(el[src_e] + er[dst_e], 0.2)
    a = _ref_softmax(e, dst_e, num_dst)
    rst = torch.zeros(num_dst, H, F, dtype=torch.float64).index_add_(0, dst_e, a[:, :, None] * fs[src_e])
    indeg = torch.bincount(dst_e, minlength=num_dst).double().clamp(min=1)
    rst = rst * indeg.pow(0.5).view(-1, 1, 1)
    rst = rst + (xd @ P["res_fc.weight"].T).view(num_dst, -1, F)
    assert float((out.detach().cpu().double() - rst.detach()).abs().max() / rst.abs().max()) < 2e-5
    w = torch.randn(num_dst, H, F, dtype=torch.float64, generator=torch.Generator().manual_seed(3))
    (rst * w).sum().backward()
    (out * w.float().to(DEV)).sum().backward()
    for k, v in conv.named_parameters():
        want = P[k].grad
        assert float((v.grad.cpu().double() - want).abs().max() / want.abs().max()) < 1e-4, k


def test_gat_model_full_graph_step(ttg_lib):
    """The reference's GAT stack (gnn_model.py:444-497) on a full graph over TT-reconstructed
    features: one training step runs and lowers the loss on a fixed batch."""
    import gnn_ops
    from FBTT.tt_embeddings_ops import OptimType, TTEmbeddingBag
    rng = np.random.default_rng(9)
    n = 2744
    blk, _, _ = _block(rng, n, n, 8)
    torch.manual_seed(1)
    emb = TTEmbeddingBag(n, 128, [16, 16], [14, 14, 14], [4, 4, 8], optimizer=OptimType.SGD,
                         learning_rate=0.05, sparse=True, use_cache=False, weight_dist="normal")
    with torch.no_grad():
        for c in emb.tt_cores:
            c.mul_(20.0)
    model = gnn_ops.GAT(128, 7, 32, 3, 3, torch.relu, dropout=0.0, norm="both").to(DEV)
    opt = torch.optim.Adam(model.parameters(), lr=0.01)
    labels = torch.randint(0, 7, (n,), device=DEV)
    ids = torch.arange(n, device=DEV)
    offs = torch.arange(n + 1, device=DEV)
    losses = []
    for _ in range(8):
        x = emb(ids, offs)
        loss = torch.nn.functional.cross_entropy(model(blk, x), labels)
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert np.isfinite(losses).all() and losses[-1] < losses[0]


@pytest.mark.parametrize("mean", [True, False])
@pytest.mark.parametrize("weighted", [False, True])
def test_aggregate_backward_as_a_gather_over_the_transposed_block(ttg_lib, mean, weighted, monkeypatch):
    """aggregate()'s backward on a full-graph (square) block runs as an SpMM over the transposed block; same
    gradient as the atomics path and as fp64 torch, with and without edge weights, rows without out-edges zero."""
    import gnn_ops
    rng = np.random.default_rng(11)
    n = 700
    blk, dst_e, src_e = _block(rng, n, n, 9)
    E = src_e.numel()
    g = torch.Generator().manual_seed(3)
    x64 = torch.randn(n, 48, generator=g, dtype=torch.float64).requires_grad_(True)
    w64 = torch.rand(E, generator=g, dtype=torch.float64) + 0.5 if weighted else None
    msg = x64[src_e] * (w64[:, None] if weighted else 1.0)
    out64 = torch.zeros(n, 48, dtype=torch.float64).index_add_(0, dst_e, msg)
    if mean:
        deg = torch.bincount(dst_e, minlength=n).clamp(min=1).double()
        out64 = out64 / deg[:, None]
    up = torch.randn(n, 48, generator=g, dtype=torch.float64)
    (out64 * up).sum().backward()
    grads = []
    for gather in (True, False):
        monkeypatch.setattr(gnn_ops, "GATHER_BACKWARD", gather)
        x = x64.detach().float().to(DEV).requires_grad_(True)
        ew = w64.float().to(DEV) if weighted else None
        out = gnn_ops.aggregate(blk, x, mean=mean, edge_weight=ew)
        assert float((out.detach().cpu().double() - out64.detach()).abs().max() / out64.abs().max()) < 1e-5
        (out * up.float().to(DEV)).sum().backward()
        grads.append(x.grad.cpu().double())
        assert float((grads[-1] - x64.grad).abs().max() / x64.grad.abs().max()) < 1e-5
    assert getattr(blk, "_transposed", None) is not None
    assert float((grads[0] - grads[1]).abs().max() / grads[1].abs().max()) < 1e-6


def test_edge_add_uv_backward_on_a_full_graph_block(ttg_lib):
    """edge_add_uv (u_add_v of the attention scores) with its gather backward against plain torch indexing."""
    import gnn_ops
    rng = np.random.default_rng(21)
    n, H = 600, 3
    blk, dst_e, src_e = _block(rng, n, n, 11)
    g = torch.Generator().manual_seed(5)
    el = torch.randn(n, H, generator=g).to(DEV).requires_grad_(True)
    er = torch.randn(n, H, generator=g).to(DEV).requires_grad_(True)
    up = torch.randn(src_e.numel(), H, generator=g).to(DEV)
    assert gnn_ops._gather_backward(blk)
    (gnn_ops.edge_add_uv(blk, el, er) * up).sum().backward()
    got = (el.grad.clone(), er.grad.clone())
    el.grad = er.grad = None
    ((el[src_e.to(DEV)] + er[dst_e.to(DEV)]) * up).sum().backward()
    for a, b in zip(got, (el.grad, er.grad)):
        assert float((a - b).abs().max() / b.abs().max()) < 1e-5
    assert float(got[1][0].abs().max()) == 0.0       # destination 0 has no in-edges
