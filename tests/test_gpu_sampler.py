"""Neighbour sampler + block builder: CUDA through the C ABI vs the oracle, bit-exact
(integer work), plus the structural properties SAGEConv relies on."""
import numpy as np
import pytest
import torch

from oracle import sampler_oracle as so

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _graph(rng, n, mean_deg):
    deg = rng.poisson(mean_deg, size=n)
    deg[:5] = [0, 1, 3, 40, 200]                 # empty, below, at and far above typical fanouts
    indptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int64)
    indices = rng.integers(0, n, size=int(indptr[-1])).astype(np.int32)
    return indptr, indices


@pytest.mark.parametrize("fanout", [1, 5, 15])
def test_block_matches_oracle_bit_exact(ttg_lib, fanout):
    import sampler
    rng = np.random.default_rng(fanout)
    n = 3000
    indptr, indices = _graph(rng, n, 12)
    g = sampler.CSRGraph(torch.from_numpy(indptr).to(DEV), torch.from_numpy(indices).to(DEV))
    dst = np.concatenate([[0, 1, 2, 3, 4], rng.permutation(np.arange(5, n))[:700]]).astype(np.int64)
    blk, src = sampler.sample_block(g, torch.from_numpy(dst).to(DEV), fanout, seed=12345)
    w_indptr, w_indices, w_src = so.sample_block(indptr, indices, dst, fanout, 12345)
    assert np.array_equal(blk.indptr.cpu().numpy(), w_indptr)
    assert np.array_equal(blk.indices.cpu().numpy(), w_indices)
    assert np.array_equal(src.cpu().numpy(), w_src)
    assert blk.num_dst == dst.size and blk.num_src == w_src.size


def test_block_properties_and_determinism(ttg_lib):
    import sampler
    rng = np.random.default_rng(0)
    n = 50000
    indptr, indices = _graph(rng, n, 30)
    g = sampler.CSRGraph(torch.from_numpy(indptr).to(DEV), torch.from_numpy(indices).to(DEV))
    seeds = torch.from_numpy(rng.permutation(n)[:1024].astype(np.int64)).to(DEV)
    s = sampler.NeighborSampler([5, 10, 15])
    inp, outp, blocks = s.sample_blocks(g, seeds, seed=7)
    inp2, _, blocks2 = s.sample_blocks(g, seeds, seed=7)
    assert torch.equal(inp, inp2) and all(torch.equal(a.indices, b.indices) for a, b in zip(blocks, blocks2))
    _, _, blocks3 = s.sample_blocks(g, seeds, seed=8)
    assert not torch.equal(blocks[2].indices, blocks3[2].indices)
    assert torch.equal(outp, seeds) and len(blocks) == 3
    assert blocks[2].num_dst == 1024 and blocks[1].num_dst == blocks[2].num_src
    assert blocks[0].num_dst == blocks[1].num_src and inp.numel() == blocks[0].num_src
    # last layer: every sampled edge is a real in-edge, counts = min(degree, fanout), no repeats
    blk = blocks[2]
    ip, ix = blk.indptr.cpu().numpy(), blk.indices.cpu().numpy()
    # source ids of the last block, global: dst nodes first
    _, src = sampler.sample_block(g, seeds, 15, 7 * 1000003 + 2)
    src = src.cpu().numpy()
    assert np.array_equal(src[:1024], seeds.cpu().numpy()) and np.unique(src).size == src.size
    for i, v in enumerate(seeds.cpu().numpy()[:200]):
        nb = indices[indptr[v]:indptr[v + 1]]
        got = src[ix[ip[i]:ip[i + 1]]]
        assert got.size == min(nb.size, 15)
        pos_ok = np.isin(got, nb).all()
        assert pos_ok
        if nb.size > 15 and np.unique(nb).size == nb.size:
            assert np.unique(got).size == got.size


def test_sampled_blocks_drive_sageconv(ttg_lib):
    import gnn_ops
    import sampler
    rng = np.random.default_rng(3)
    n = 20000
    indptr, indices = _graph(rng, n, 20)
    g = sampler.CSRGraph(torch.from_numpy(indptr).to(DEV), torch.from_numpy(indices).to(DEV))
    seeds = torch.from_numpy(rng.permutation(n)[:256].astype(np.int64)).to(DEV)
    inp, _, blocks = sampler.NeighborSampler([5, 10]).sample_blocks(g, seeds, seed=1)
    torch.manual_seed(0)
    x = torch.randn(inp.numel(), 32, device=DEV)
    l0, l1 = gnn_ops.SAGEConv(32, 64).to(DEV), gnn_ops.SAGEConv(64, 8).to(DEV)
    h = torch.relu(l0(blocks[0], (x, x[:blocks[0].num_dst])))
    out = l1(blocks[1], (h, h[:blocks[1].num_dst]))
    assert out.shape == (256, 8) and torch.isfinite(out).all()
    # dense reference of the first layer's mean aggregation
    b = blocks[0]
    ip, ix = b.indptr.cpu().numpy(), b.indices.cpu().numpy().astype(np.int64)
    xc = x.cpu().numpy()
    want = np.zeros((b.num_dst, 32), np.float32)
    for v in range(b.num_dst):
        if ip[v + 1] > ip[v]:
            want[v] = xc[ix[ip[v]:ip[v + 1]]].mean(0)
    got = gnn_ops.aggregate(b, x, mean=True).cpu().numpy()
    assert np.abs(got - want).max() < 1e-5


@pytest.mark.parametrize("thread", [False, True], ids=["same-thread", "worker-thread"])
def test_prefetched_minibatches_equal_direct_sampling(ttg_lib, thread):
    """Sampling ahead on a side stream (from the consumer's thread or from the worker thread) yields
    exactly the batches of the plain loop, and the tensors are safe to consume on the main stream
    while the next samples are being drawn."""
    import sage
    import sampler
    g = sage.synthetic_graph(20000, 400000, torch.device(DEV), seed=3)
    smp = sampler.NeighborSampler([3, 5])
    seeds_all = torch.randperm(20000, generator=torch.Generator().manual_seed(1)).to(DEV)
    seeds_of = lambda s: seeds_all[s * 256:(s + 1) * 256]
    direct = [smp.sample_blocks(g, seeds_of(s), seed=100 + s) for s in range(6)]
    sums = []
    n = 0
    for inp, outp, blocks in sampler.prefetched_minibatches(g, smp, seeds_of, lambda s: 100 + s, 6, thread=thread):
        torch.cuda._sleep(3_000_000)                  # a slow "training step" on the main stream
        sums.append((inp.sum(), blocks[0].indices.long().sum()))
        d_inp, d_outp, d_blocks = direct[n]
        assert torch.equal(inp, d_inp) and torch.equal(outp, d_outp)
        for b, d in zip(blocks, d_blocks):
            assert torch.equal(b.indptr, d.indptr) and torch.equal(b.indices, d.indices)
            assert (b.num_src, b.num_dst) == (d.num_src, d.num_dst)
        n += 1
    assert n == 6
    torch.cuda.synchronize()
    for (a, b), (d_inp, _, d_blocks) in zip(sums, direct):
        assert int(a) == int(d_inp.sum()) and int(b) == int(d_blocks[0].indices.long().sum())
    assert list(sampler.prefetched_minibatches(g, smp, seeds_of, lambda s: s, 0, thread=thread)) == []


def test_prefetch_worker_hands_errors_over_and_stops_when_the_consumer_leaves(ttg_lib):
    import threading
    import sage
    import sampler
    g = sage.synthetic_graph(5000, 60000, torch.device(DEV), seed=4)
    smp = sampler.NeighborSampler([3, 3])
    seeds = torch.arange(128, device=DEV)

    def bad_seeds(s):
        if s == 2:
            raise ValueError("no seeds for step 2")
        return seeds

    got = 0
    with pytest.raises(ValueError, match="step 2"):
        for _ in sampler.prefetched_minibatches(g, smp, bad_seeds, lambda s: s, 5, thread=True):
            got += 1
    assert got == 2
    it = sampler.prefetched_minibatches(g, smp, lambda s: seeds, lambda s: s, 50, thread=True)
    next(it)
    it.close()                                     # the consumer stops early: the worker must not linger
    assert not [t for t in threading.enumerate() if t.name == "ttg-sampler" and t.is_alive()]
