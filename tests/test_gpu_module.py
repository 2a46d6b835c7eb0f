"""GPU tests of the module surface, written the way the reference's own (commented-out) unit
test states correctness (sage_profiler.py:303-305, 362-367, 405-426, 466-500):
forward == sum-mode EmbeddingBag over full_weight(); dense grads == autograd through
tt_matrix_to_full; fused SGD == core - lr * grad; Adagrad == state = g^2, core - lr*g/(sqrt+eps).
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _bags(rng, n_emb, n_bags, one_per_bag):
    lengths = (np.ones(n_bags, dtype=np.int64) if one_per_bag
               else np.clip(np.round(rng.normal(4, 5, size=n_bags)), 0, None).astype(np.int64))
    offsets = np.concatenate([[0], np.cumsum(lengths)]).astype(np.int64)
    indices = rng.integers(0, n_emb, size=int(offsets[-1])).astype(np.int64)
    return torch.from_numpy(indices).to(DEV), torch.from_numpy(offsets).to(DEV)


def _make(n_emb, D, ranks, p, q, **kw):
    from FBTT.tt_embeddings_ops import TTEmbeddingBag
    torch.manual_seed(0)
    np.random.seed(0)
    kw.setdefault("weight_dist", "normal")
    kw.setdefault("use_cache", False)
    m = TTEmbeddingBag(n_emb, D, ranks, p, q, **kw)
    with torch.no_grad():
        for c in m.tt_cores:     # O(1) entries so that 1e-5 relative is a meaningful bar
            c.mul_(n_emb ** 0.5 * 0.5)
    return m


@pytest.mark.parametrize("one_per_bag", [True, False])
@pytest.mark.parametrize("cfg", [
    (2708, 128, [16, 16], [14, 14, 14], [4, 4, 8]),
    (5 * 6 * 7, 100, [16, 16], [5, 6, 7], [4, 5, 5]),
    (99, 32, [12], [9, 11], [4, 8]),
    (360, 48, [4, 6, 5], [3, 4, 5, 6], [2, 2, 3, 4]),
])
def test_forward_equals_embedding_bag_and_dense_grads_equal_autograd(ttg_lib, cfg, one_per_bag):
    from FBTT.tt_embeddings_ops import tt_matrix_to_full
    n_emb, D, ranks, p, q = cfg
    m = _make(n_emb, D, ranks, p, q, sparse=False)
    rng = np.random.default_rng(1)
    indices, offsets = _bags(rng, n_emb, 257, one_per_bag)
    out = m(indices.int(), offsets.int())          # DGL hands int32 ids (sage_dgl_partition.py:85)
    W = m.full_weight()
    ref = torch.nn.functional.embedding_bag(indices, W, offsets, mode="sum", include_last_offset=True)
    assert out.shape == ref.shape
    # 1e-5 relative to the scale of the table (bag sums cancel, so not elementwise-relative)
    torch.testing.assert_close(out, ref, rtol=1e-5, atol=1e-5 * float(ref.abs().max()))
    d_out = torch.rand_like(out) * 0.1
    out.backward(d_out)
    cores64 = [c.detach().double().requires_grad_(True) for c in m.tt_cores]
    W64 = tt_matrix_to_full(p, q, [1] + ranks + [1], cores64, [1, 0, 2, 3]).double()
    ref64 = torch.nn.functional.embedding_bag(indices, W64, offsets, mode="sum",
                                              include_last_offset=True)
    (ref64 * d_out.double()).sum().backward()
    for c, c64 in zip(m.tt_cores, cores64):
        torch.testing.assert_close(c.grad, c64.grad.float(), rtol=1e-5,
                                   atol=1e-5 * float(c64.grad.abs().max()))


def test_sparse_sgd_updates_cores_in_place(ttg_lib):
    from FBTT.tt_embeddings_ops import OptimType
    n_emb, D, ranks, p, q = 5 * 6 * 7, 100, [16, 16], [5, 6, 7], [4, 5, 5]
    m = _make(n_emb, D, ranks, p, q, sparse=True, optimizer=OptimType.SGD, learning_rate=0.1)
    m_dense = _make(n_emb, D, ranks, p, q, sparse=False)
    rng = np.random.default_rng(2)
    indices, offsets = _bags(rng, n_emb, 128, False)
    before = [c.detach().clone() for c in m.tt_cores]
    out = m(indices, offsets)
    d_out = torch.rand_like(out) * 0.1
    out.backward(d_out)
    assert all(c.grad is None for c in m.tt_cores)      # fused path returns no gradients
    out_d = m_dense(indices, offsets)
    out_d.backward(d_out)
    for b, c, cd in zip(before, m.tt_cores, m_dense.tt_cores):
        torch.testing.assert_close(c.detach(), b - 0.1 * cd.grad, rtol=1e-5, atol=1e-6)


def test_cache_flow_matches_uncached_until_rows_are_trained(ttg_lib):
    n_emb, D, ranks, p, q = 5 * 6 * 7, 100, [16, 16], [5, 6, 7], [4, 5, 5]
    m = _make(n_emb, D, ranks, p, q, sparse=False, use_cache=True, cache_size=20,
              hashtbl_size=4001)
    rng = np.random.default_rng(3)
    hot = torch.from_numpy(rng.choice(n_emb, size=15, replace=False)).to(DEV)
    for _ in range(3):                                    # warm-up epoch: statistics only
        idx = torch.cat([hot, torch.from_numpy(rng.integers(0, n_emb, size=30)).to(DEV)])
        off = torch.arange(idx.numel() + 1, device=DEV)
        assert m.warmup
        m(idx, off)
    m.cache_populate()
    assert not m.warmup
    idx = torch.cat([hot[:7], torch.from_numpy(rng.integers(0, n_emb, size=50)).to(DEV), hot[7:]])
    off = torch.arange(idx.numel() + 1, device=DEV)
    out = m(idx, off)
    ref = m.full_weight()[idx]
    torch.testing.assert_close(out, ref, rtol=1e-5, atol=1e-6)   # cache rows are a snapshot
    out.backward(torch.ones_like(out))
    assert m.cache_weight.grad is not None and float(m.cache_weight.grad.abs().sum()) > 0
    # the hot rows were served from the cache: their gradient went to cache_weight, not the cores
    loc_rows = (m.cache_weight.grad.abs().sum(dim=1) > 0).sum().item()
    assert loc_rows >= 10
    sd = m.state_dict()
    for k in ["L", "hashtbl", "cache_freq", "cache_state", "cache_weight", "tt_cores.0"]:
        assert k in sd


def test_merge_lfu_statistics_rebuilds_the_table_with_the_same_counts(ttg_lib):
    """dp.merge_lfu_statistics on one rank is the identity on (key -> count): the table is rebuilt from the merged
    pairs (the two-rank merge itself is covered on CPU, tests/test_dp_gloo_cpu.py) and cache_populate() then
    caches the most frequent rows."""
    import dp
    n_emb, D, ranks, p, q = 5 * 6 * 7, 100, [16, 16], [5, 6, 7], [4, 5, 5]
    m = _make(n_emb, D, ranks, p, q, sparse=False, use_cache=True, cache_size=20, hashtbl_size=4001)
    rng = np.random.default_rng(5)
    hot = torch.from_numpy(rng.choice(n_emb, size=12, replace=False)).to(DEV)
    for _ in range(4):
        idx = torch.cat([hot, torch.from_numpy(rng.integers(0, n_emb, size=40)).to(DEV)])
        m(idx, torch.arange(idx.numel() + 1, device=DEV))

    def table(mod):
        used = mod.hashtbl >= 0
        return dict(zip(mod.hashtbl[used].tolist(), mod.cache_freq[used].tolist()))

    before = table(m)
    assert len(before) > 12 and all(before[int(h)] >= 4 for h in hot.tolist())
    assert dp.merge_lfu_statistics(m) == len(before)
    assert table(m) == before
    m.cache_populate()
    cached = set(m.hashtbl[m.cache_state >= 0].tolist())
    assert set(hot.tolist()) <= cached and len(cached) <= 20


def test_eff_embedding_forward_and_fused_sgd(ttg_lib):
    from Efficient_TT.efficient_tt import Eff_TTEmbedding
    from FBTT.tt_embeddings_ops import tt_matrix_to_full
    p, q, ranks = [5, 6, 7], [4, 5, 5], [16, 16]
    n_emb = 5 * 6 * 7
    torch.manual_seed(0)
    m = Eff_TTEmbedding(n_emb, 100, ranks, p, q, learning_rate=0.1, device=0)
    with torch.no_grad():
        for c in m.tt_cores:
            c.mul_(20.0)
    rng = np.random.default_rng(4)
    idx = torch.from_numpy(rng.integers(0, n_emb, size=500)).to(DEV)
    idx[10] = idx[11]
    cores0 = [c.detach().clone() for c in m.tt_cores]
    out = m(idx)
    W = tt_matrix_to_full(p, q, [1] + ranks + [1], [c[None] for c in cores0], [1, 0, 2, 3])
    torch.testing.assert_close(out, W[idx], rtol=1e-5, atol=1e-6)
    d_out = torch.rand_like(out) * 0.1
    out.backward(d_out)
    c64 = [c.double()[None].requires_grad_(True) for c in cores0]
    W64 = tt_matrix_to_full(p, q, [1] + ranks + [1], c64, [1, 0, 2, 3]).double()
    (W64[idx] * d_out.double()).sum().backward()
    for c, c0, g in zip(m.tt_cores, cores0, c64):
        torch.testing.assert_close(c.detach(), c0 - 0.1 * g.grad[0].float(), rtol=1e-5, atol=1e-6)


def test_sageconv_and_graphconv_against_dense_torch(ttg_lib):
    import gnn_ops
    from helpers import random_block
    rng = np.random.default_rng(6)
    num_src, num_dst = 900, 300
    indptr, indices = random_block(rng, num_src, num_dst, 9)
    blk = gnn_ops.Block(torch.from_numpy(indptr).to(DEV), torch.from_numpy(indices).to(DEV),
                        num_src, num_dst)
    A = torch.zeros(num_dst, num_src, dtype=torch.float64)
    for v in range(num_dst):
        for e in range(indptr[v], indptr[v + 1]):
            A[v, indices[e]] += 1
    A = A.to(DEV)
    deg = A.sum(1).clamp(min=1)
    # aggregate-then-linear / linear-then-aggregate; tuple input as the reference passes it (one
    # fused autograd node for both reads when h_dst is the head of h), a separate h_dst, plain h
    for fin, fout, how in [(100, 256, "head"), (256, 47, "head"), (100, 256, "plain"), (100, 256, "copy")]:
        torch.manual_seed(1)
        conv = gnn_ops.SAGEConv(fin, fout, "mean").to(DEV)
        h = torch.randn(num_src, fin, device=DEV, requires_grad=True)
        out = conv(blk, {"head": (h, h[:num_dst]), "plain": h, "copy": (h, h[:num_dst] * 1.0)}[how])
        h64 = h.detach().double().requires_grad_(True)
        neigh = (A @ h64) / deg[:, None]
        ref = (h64[:num_dst] @ conv.fc_self.weight.double().t()
               + neigh @ conv.fc_neigh.weight.double().t() + conv.bias.double())
        torch.testing.assert_close(out, ref.float(), rtol=1e-4, atol=1e-4)
        g = torch.randn_like(out)
        out.backward(g)
        ref.backward(g.double())
        torch.testing.assert_close(h.grad, h64.grad.float(), rtol=1e-4, atol=1e-4)
    conv = gnn_ops.GraphConv(64, 32).to(DEV)
    h = torch.randn(num_src, 64, device=DEV)
    out = conv(blk, h)
    odeg = A.sum(0).clamp(min=1)
    ref = ((A @ (h.double() / odeg.sqrt()[:, None])) / deg.sqrt()[:, None]) @ conv.weight.double() \
        + conv.bias.double()
    torch.testing.assert_close(out, ref.float(), rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("cfg", [
    (169343, 128, [16, 16], [55, 55, 56], [4, 4, 8]),
    (2708, 128, [16, 16], [14, 14, 14], [4, 4, 8]),
    (60000, 100, [16, 16], [30, 40, 50], [4, 5, 5]),
])
def test_rows_range_equals_tt_matrix_to_full(ttg_lib, cfg):
    """the plan-free range reconstruction against the reference's pure-PyTorch
    contraction (FBTT/tt_embeddings_ops.py:80-127), whole table and an unaligned slice."""
    from FBTT.tt_embeddings_ops import tt_matrix_to_full
    n_emb, D, ranks, p, q = cfg
    m = _make(n_emb, D, ranks, p, q, sparse=False)
    W = tt_matrix_to_full(p, q, [1] + ranks + [1], [c.detach() for c in m.tt_cores], [1, 0, 2, 3])
    with torch.no_grad():
        full = m.full_weight()
    assert full.shape == W.shape
    assert float((full - W).abs().max() / W.abs().max()) < 1e-5
    a, n = 1237, 20011 if n_emb > 30000 else 997
    part = m.rows_range(a, n)
    assert torch.equal(part, full[a:a + n])
    # and the same rows through the indexed forward
    idx = torch.arange(a, a + n, device=full.device)
    out = m(idx, torch.arange(n + 1, device=full.device))
    assert float((out - part).abs().max() / W.abs().max()) < 1e-6


@pytest.mark.gpu
def test_host_batch_pipeline_and_deferred_scalars():
    """Batches staged a step ahead arrive intact and in order; a slot is not overwritten while
    the step that reads it is still running; deferred scalars come back in push order."""
    import pipeline
    dev = torch.device("cuda", 0)
    n = 1 << 20
    host = [torch.full((n,), i, dtype=torch.int64).pin_memory() for i in range(6)]
    offs = torch.arange(n + 1, dtype=torch.int64).pin_memory()
    pipe = pipeline.HostBatchPipeline(dev, depth=2)
    reader = pipeline.DeferredScalars(dev, delay=1)
    got = []
    pipe.put(host[0], offs)
    with pytest.raises(RuntimeError):
        pipe.put(torch.zeros(4, dtype=torch.int64))          # not pinned
    for i in range(6):
        idx, off = pipe.get()
        if i + 1 < 6:
            pipe.put(host[i + 1], offs)
        torch.cuda._sleep(2_000_000)                          # the step is slow; the prefetch is not
        s = (idx.double().mean() + off[-1].double() * 0).float()
        pipe.release()
        got += reader.push(s)
    got += reader.drain()
    assert got == [float(i) for i in range(6)]
    assert pipe.h2d_bytes == 6 * (n * 8 + (n + 1) * 8) and reader.d2h_bytes == 6 * 4
    with pytest.raises(RuntimeError):
        pipe.get()


def test_graphed_step_equals_eager_steps(ttg_lib):
    """A TTEmbeddingBag step (forward, loss.backward with the fused SGD update) replayed as a CUDA
    graph on refilled index buffers leaves the same cores as the same steps launched eagerly."""
    import pipeline
    from FBTT.tt_embeddings_ops import OptimType
    n_emb, D, ranks, p, q = 2708, 128, [16, 16], [14, 14, 14], [4, 4, 8]
    kw = dict(sparse=True, optimizer=OptimType.SGD, learning_rate=0.05)
    m_graph, m_eager = _make(n_emb, D, ranks, p, q, **kw), _make(n_emb, D, ranks, p, q, **kw)
    for a, b in zip(m_graph.tt_cores, m_eager.tt_cores):
        assert torch.equal(a, b)
    rng = np.random.default_rng(5)
    nb = 300
    batches = [torch.from_numpy(rng.integers(0, n_emb, size=nb)).to(DEV) for _ in range(4)]
    offsets = torch.arange(nb + 1, device=DEV)
    target = torch.rand(nb, D, device=DEV) * 0.01

    def step(m, idx):
        loss = (m(idx, offsets) * target).sum()
        loss.backward()
        return loss

    idx_static = batches[0].clone()
    gs = pipeline.GraphedStep(lambda: step(m_graph, idx_static), torch.device(DEV), warmup=2)
    eager_losses = [float(step(m_eager, batches[0])) for _ in range(2)]
    graph_losses = []
    for b in batches[1:]:
        idx_static.copy_(b)
        graph_losses.append(float(gs()))
        eager_losses.append(float(step(m_eager, b)))
    np.testing.assert_allclose(graph_losses, eager_losses[2:], rtol=1e-4)
    for a, b in zip(m_graph.tt_cores, m_eager.tt_cores):
        torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-5 * float(b.abs().max()))
    assert not torch.equal(m_eager.tt_cores[2], _make(n_emb, D, ranks, p, q, **kw).tt_cores[2])


def test_gcn_stack_wiring(ttg_lib):
    """gnn_ops.GCN composes GraphConv / Linear / BatchNorm as the reference's GCN does
    (gnn_model.py:297-314): checked against the same layers applied by hand, in eval mode."""
    import gnn_ops
    from helpers import random_block
    rng = np.random.default_rng(8)
    n = 400
    indptr, indices = random_block(rng, n, n, 6)
    blk = gnn_ops.Block(torch.from_numpy(indptr).to(DEV), torch.from_numpy(indices).to(DEV), n, n)
    torch.manual_seed(2)
    for use_linear in (False, True):
        m = gnn_ops.GCN(32, 24, 5, 3, torch.nn.functional.relu, 0.5, use_linear).to(DEV).eval()
        x = torch.randn(n, 32, device=DEV, requires_grad=True)
        out = m(blk, x)
        h = x
        for i in range(3):
            c = m.convs[i](blk, h)
            h = c + m.linear[i](h) if use_linear else c
            if i < 2:
                h = torch.relu(m.bns[i](h))
        torch.testing.assert_close(out, h)
        assert out.shape == (n, 5) and m.convs[2].bias is not None and m.convs[0].bias is None
        out.sum().backward()
        assert x.grad is not None and float(x.grad.abs().sum()) > 0


@pytest.mark.parametrize("sparse", [False, True])
def test_device_side_cache_split_equals_the_partitioned_flow(ttg_lib, sparse):
    """The module's split without a host count (tt_embeddings.cache_mark: cached entries become id -1 for the TT
    kernels, uncached ones location -1 for the cache kernels) against the reference's data flow (partitioned
    lists, preprocess_indices_sync): same output, same gradients / same fused updates, bags partly cached."""
    n_emb, D, ranks, p, q = 5 * 6 * 7, 100, [16, 16], [5, 6, 7], [4, 5, 5]
    rng = np.random.default_rng(9)
    hot = torch.from_numpy(rng.choice(n_emb, size=15, replace=False)).to(DEV)
    lengths = rng.integers(0, 5, size=60)
    idx = torch.from_numpy(np.concatenate([hot.cpu().numpy(), rng.integers(0, n_emb, size=int(lengths.sum()) - 15)])).to(DEV)
    idx = idx[torch.randperm(idx.numel(), generator=torch.Generator().manual_seed(1)).to(DEV)]
    off = torch.from_numpy(np.concatenate([[0], np.cumsum(lengths)])).to(DEV)
    res = []
    for on_device in (True, False):
        m = _make(n_emb, D, ranks, p, q, sparse=sparse, use_cache=True, cache_size=20, hashtbl_size=4001,
                  learning_rate=0.1)
        m.split_on_device = on_device
        for _ in range(3):
            w = torch.cat([hot, torch.from_numpy(rng.integers(0, n_emb, size=30)).to(DEV)])
            m(w, torch.arange(w.numel() + 1, device=DEV))
        m.cache_populate()
        out = m(idx, off)
        g = torch.Generator().manual_seed(2)
        out.backward((torch.rand(out.shape, generator=g) * 0.1).to(DEV))
        torch.cuda.synchronize()
        rec = [out.detach().clone()]
        if sparse:
            rec += [c.detach().clone() for c in m.tt_cores] + [m.cache_weight.detach().clone()]
        else:
            rec += [c.grad.clone() for c in m.tt_cores] + [m.cache_weight.grad.clone()]
        res.append(rec)
        rng = np.random.default_rng(9)       # the same warm-up stream for the second module
        rng.choice(n_emb, size=15, replace=False)
        rng.integers(0, 5, size=60)
        rng.integers(0, n_emb, size=int(lengths.sum()) - 15)
    assert float(res[0][-1].abs().sum()) > 0
    for a, b in zip(res[0], res[1]):
        torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-6)


def test_cached_module_step_is_capturable(ttg_lib):
    """With the split on the device nothing in forward + backward synchronises the stream: the cached module's
    step runs as a CUDA graph and replays to the eager result."""
    n_emb, D, ranks, p, q = 5 * 6 * 7, 100, [16, 16], [5, 6, 7], [4, 5, 5]
    rng = np.random.default_rng(11)
    hot = torch.from_numpy(rng.choice(n_emb, size=15, replace=False)).to(DEV)
    m = _make(n_emb, D, ranks, p, q, sparse=False, use_cache=True, cache_size=20, hashtbl_size=4001)
    for _ in range(3):
        w = torch.cat([hot, torch.from_numpy(rng.integers(0, n_emb, size=30)).to(DEV)])
        m(w, torch.arange(w.numel() + 1, device=DEV))
    m.cache_populate()
    idx = torch.cat([hot[:8], torch.from_numpy(rng.integers(0, n_emb, size=40)).to(DEV)])
    off = torch.arange(idx.numel() + 1, device=DEV)
    eager = m(idx, off).detach().clone()
    torch.cuda.synchronize()
    static_out = torch.empty_like(eager)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=s):
            static_out.copy_(m(idx, off).detach())
    torch.cuda.current_stream().wait_stream(s)
    static_out.zero_()
    gr.replay()
    torch.cuda.synchronize()
    torch.testing.assert_close(static_out, eager, rtol=1e-6, atol=1e-7)
    m.split_on_device = False
    with pytest.raises(RuntimeError):
        with torch.cuda.stream(s):
            g2 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g2, stream=s):
                m(idx, off)
