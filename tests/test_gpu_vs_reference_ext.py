"""Parity of the CUDA path against the REFERENCE'S OWN CUDA KERNELS, built unmodified for sm_100a
by oracle/build_ref.py (FBTT/tt_embeddings.cpp + tt_embeddings_cuda.cu, shipped to the GPU box
as oracle/_ref/*.so).  Skipped when that build is absent.

Tolerances.  OUR results are held to 1e-5 of the tensor's largest magnitude against the fp64 oracle
(the truth; north star).  The reference's own kernels are fp32 cuBLAS GEMMs plus float atomicAdd
scatters in no fixed order: their distance from the fp64 truth is measured in the same test and is the
stated bound of the direct ours-vs-reference comparison (REF_BOUND, a few 1e-5 on gradients that sum
thousands of rows per element).  Index path: bit-exact."""
import numpy as np
import pytest
import torch

from oracle import ref_ext

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-5
REF_BOUND = 5e-5     # reference kernels vs fp64 truth on the gradients (float atomics); asserted below

SHAPES = {
    "products": ([125, 140, 140], [4, 5, 5], [1, 16, 16, 1], 2449029),
    "arxiv": ([55, 55, 56], [4, 4, 8], [1, 16, 16, 1], 169343),
    "papers": ([481, 481, 481], [4, 4, 8], [1, 32, 32, 1], 111059956),
}


@pytest.fixture(scope="module")
def ref():
    m = ref_ext.load()
    if m is None:
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    return m


def _cores(p, q, r, seed):
    g = torch.Generator().manual_seed(seed)
    return [(torch.randn(1, p[t], r[t] * q[t] * r[t + 1], generator=g) * (0.5 / np.sqrt(r[t]))).to(DEV)
            for t in range(3)]


def _L(p):
    return torch.tensor([p[1] * p[2], p[2], 1], dtype=torch.int64, device=DEV)


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


@pytest.mark.parametrize("shape,nnz", [("products", 40000), ("arxiv", 20000), ("papers", 6000),
                                       ("products", 3000)])
def test_forward_and_dense_backward_match_reference_kernels(ttg_lib, ref, shape, nnz):
    import tt_embeddings as te
    p, q, r, n_emb = SHAPES[shape]
    D = int(np.prod(q))
    cores = _cores(p, q, r, 5)
    rng = np.random.default_rng(1)
    hi = min(n_emb, 90000) if nnz > 10000 else n_emb
    idx = torch.from_numpy(rng.integers(0, hi, size=nnz).astype(np.int64)).to(DEV)
    row = torch.arange(nnz, device=DEV)
    tb = torch.zeros(nnz, dtype=torch.int64, device=DEV)
    L = _L(p)
    want = ref.tt_forward(1000, 1, nnz, D, p, q, r, L, nnz, idx, row, tb, cores)
    got = te.tt_forward(1000, 1, nnz, D, p, q, r, L, nnz, idx, row, tb, cores)
    assert _rel(got, want) < TOL
    dO = (torch.rand(1, nnz, D, generator=torch.Generator().manual_seed(2)) * 0.1).to(DEV)
    wd = ref.tt_dense_backward(1000, D, p, q, r, L, nnz, idx, row, tb, dO, cores)
    gd = te.tt_dense_backward(1000, D, p, q, r, L, nnz, idx, row, tb, dO, cores)
    from oracle import oracle as orc
    truth = orc.tt_backward_dense(p, q, r, [c.cpu().numpy() for c in cores], idx.cpu().numpy(),
                                  row.cpu().numpy(), dO.cpu().numpy())
    for t in range(3):
        tr = torch.from_numpy(truth[t]).to(DEV)
        assert _rel(gd[t], tr) < TOL, "core %d: ours vs fp64 oracle" % t
        assert _rel(wd[t], tr) < REF_BOUND, "core %d: reference kernels vs fp64 oracle" % t
        assert _rel(gd[t], wd[t]) < REF_BOUND, "core %d: ours vs reference kernels" % t


@pytest.mark.parametrize("shape", ["products", "arxiv", "papers"])
def test_full_size_three_way(ttg_lib, ref, shape):
    """BASELINE batch (262,144 rows; the whole table at arxiv shape) over the WHOLE index range (ids beyond 2^24
    and 2^26 at papers shape): ours vs the fp64 oracle at 1e-5, the reference's kernels vs the oracle within
    their stated bound, ours vs the reference's kernels within that bound."""
    import tt_embeddings as te
    from oracle import oracle as orc
    orc.use_all_host_threads()
    p, q, r, n_emb = SHAPES[shape]
    D = int(np.prod(q))
    nnz = min(262144, n_emb)
    cores = _cores(p, q, r, 8)
    g = torch.Generator().manual_seed(3)
    if n_emb <= (1 << 22):
        idx_c = torch.randperm(n_emb, generator=g)[:nnz]
    else:
        idx_c = torch.randint(0, n_emb, (nnz,), generator=g)
        assert int(idx_c.max()) > (1 << 26)
    idx = idx_c.to(DEV)
    row = torch.randperm(nnz, generator=g).to(DEV)        # rows in arbitrary order
    tb = torch.zeros(nnz, dtype=torch.int64, device=DEV)
    L = _L(p)
    cn = [c.cpu().numpy() for c in cores]
    want = ref.tt_forward(1000, 1, nnz, D, p, q, r, L, nnz, idx, row, tb, cores)
    got = te.tt_forward(1000, 1, nnz, D, p, q, r, L, nnz, idx, row, tb, cores)
    truth = torch.from_numpy(orc.tt_forward(p, q, r, cn, idx_c.numpy(), row.cpu().numpy(), nnz)).to(DEV)
    assert _rel(got, truth) < TOL
    assert _rel(want, truth) < TOL
    assert _rel(got, want) < TOL
    dO = ((torch.rand(1, nnz, D, generator=g) - 0.5) * 0.2).to(DEV)
    gd = te.tt_dense_backward(1000, D, p, q, r, L, nnz, idx, row, tb, dO, cores)
    wd = ref.tt_dense_backward(1000, D, p, q, r, L, nnz, idx, row, tb, dO, cores)
    td = orc.tt_backward_dense(p, q, r, cn, idx_c.numpy(), row.cpu().numpy(), dO.cpu().numpy())
    for t in range(3):
        tr = torch.from_numpy(td[t]).to(DEV)
        assert _rel(gd[t], tr) < TOL, "core %d: ours vs fp64 oracle" % t
        assert _rel(wd[t], tr) < REF_BOUND, "core %d: reference kernels vs fp64 oracle" % t
        assert _rel(gd[t], wd[t]) < REF_BOUND, "core %d: ours vs reference kernels" % t


def test_bags_with_several_indices_match_reference_kernels(ttg_lib, ref):
    import tt_embeddings as te
    p, q, r, n_emb = SHAPES["products"]
    cores = _cores(p, q, r, 6)
    rng = np.random.default_rng(3)
    lengths = rng.integers(0, 6, size=4000)
    offsets = np.concatenate([[0], np.cumsum(lengths)]).astype(np.int64)
    nnz, B = int(offsets[-1]), lengths.size
    idx = torch.from_numpy(rng.integers(0, 50000, size=nnz).astype(np.int64)).to(DEV)
    row = torch.from_numpy(np.repeat(np.arange(B), lengths).astype(np.int64)).to(DEV)
    tb = torch.zeros(nnz, dtype=torch.int64, device=DEV)
    L = _L(p)
    want = ref.tt_forward(1000, 1, B, 100, p, q, r, L, nnz, idx, row, tb, cores)
    got = te.tt_forward(1000, 1, B, 100, p, q, r, L, nnz, idx, row, tb, cores)
    assert _rel(got, want) < TOL


def test_fused_sgd_matches_reference_on_the_rows_it_updates(ttg_lib, ref):
    """the reference's launch config never updates rows >= ceil(cols/ty)*ty of a core
    with more rows than columns; on the rows it does update, the two agree."""
    import tt_embeddings as te
    from oracle import oracle as orc
    p, q, r, n_emb = SHAPES["products"]
    nnz = 30000
    rng = np.random.default_rng(4)
    idx = torch.from_numpy(rng.integers(0, n_emb, size=nnz).astype(np.int64)).to(DEV)
    row = torch.arange(nnz, device=DEV)
    tb = torch.zeros(nnz, dtype=torch.int64, device=DEV)
    dO = (torch.rand(1, nnz, 100, generator=torch.Generator().manual_seed(5)) * 0.1).to(DEV)
    L = _L(p)
    c_ref = _cores(p, q, r, 7)
    c_our = [c.clone() for c in c_ref]
    before = [c.clone() for c in c_ref]
    ref.tt_sgd_backward(1000, 100, 0.1, p, q, r, L, nnz, idx, row, tb, dO, c_ref)
    te.tt_sgd_backward(1000, 100, 0.1, p, q, r, L, nnz, idx, row, tb, dO, c_our)
    cols = [r[t] * q[t] * r[t + 1] for t in range(3)]
    lim = orc.reference_sgd_rows_updated(p, cols)
    import _ttg
    c_gen = [c.clone() for c in before]
    te.EXTRA_FLAGS = _ttg.FLAG_FORCE_GENERIC
    try:
        te.tt_sgd_backward(1000, 100, 0.1, p, q, r, L, nnz, idx, row, tb, dO, c_gen)
    finally:
        te.EXTRA_FLAGS = 0
    for t in range(3):
        n = lim[t]
        assert _rel(c_our[t][:, :n], c_ref[t][:, :n]) < TOL, (
            "core %d: ours vs reference %.3g, ours vs our generic kernels %.3g, reference vs generic %.3g"
            % (t, _rel(c_our[t][:, :n], c_ref[t][:, :n]), _rel(c_our[t], c_gen[t]),
               _rel(c_ref[t][:, :n], c_gen[t][:, :n])))
        if n < p[t]:      # the tail the reference skips: untouched there, updated here
            assert torch.equal(c_ref[t][:, n:], before[t][:, n:])
            assert not torch.equal(c_our[t][:, n:], before[t][:, n:])


def test_cache_index_path_matches_reference_kernels(ttg_lib, ref):
    """update_cache_state -> cache_populate -> preprocess_indices_sync -> cache_forward, the
    same calls on both extensions.  Insertion races make hash-table slots order dependent, so the
    state is compared as key -> (frequency, cached) maps and the partitions as sets + order rules."""
    import tt_embeddings as te
    p, q, r, n_emb = SHAPES["arxiv"]
    D = 128
    cores = _cores(p, q, r, 8)
    L = _L(p)
    size, cache_size = 20011, 500
    rng = np.random.default_rng(6)
    # keys whose primary slots are distinct and not adjacent: with colliding keys which one of
    # them is dropped after three probes depends on the insertion race in BOTH implementations
    from helpers import collision_free_keys
    from oracle import oracle as orc
    pop = collision_free_keys(orc, size, 2500, rng, 0, n_emb)
    stream = np.repeat(pop, rng.integers(1, 12, size=pop.size))
    rng.shuffle(stream)
    state = {}
    for name, ext in (("ref", ref), ("ours", te)):
        ht = torch.full((size,), -1, dtype=torch.int64, device=DEV)
        fr = torch.zeros(size, dtype=torch.int64, device=DEV)
        cs = torch.full((size,), -1, dtype=torch.int32, device=DEV)
        cw = torch.zeros(cache_size, D, device=DEV)
        ext.update_cache_state(torch.from_numpy(stream).to(DEV), ht, fr)
        ext.cache_populate(n_emb, p, q, r, cores, L, ht, fr, cs, cw)
        state[name] = (ht, fr, cs, cw)
    maps = {}
    for name, (ht, fr, cs, cw) in state.items():
        h, f, c = ht.cpu().numpy(), fr.cpu().numpy(), cs.cpu().numpy()
        w = cw.cpu().numpy()
        maps[name] = {int(k): (int(f[i]), None if c[i] < 0 else w[c[i]]) for i, k in enumerate(h) if k >= 0}
    assert set(maps["ref"]) == set(maps["ours"])
    ncached = 0
    for k, (f, wrow) in maps["ref"].items():
        f2, wrow2 = maps["ours"][k]
        assert f == f2
        if wrow is not None and wrow2 is not None:
            ncached += 1
            assert np.abs(wrow - wrow2).max() <= TOL * max(np.abs(wrow).max(), 1e-30)
    assert ncached > 0
    # lookups through each extension's own state give the same rows
    q_idx = torch.from_numpy(rng.permutation(pop)[:2000].astype(np.int64)).to(DEV)
    offs = torch.arange(q_idx.numel() + 1, device=DEV)
    outs = {}
    for name, ext in (("ref", ref), ("ours", te)):
        ht, fr, cs, cw = state[name]
        col, rowidx, tbl, n_tt, loc = ext.preprocess_indices_sync(q_idx, offs, 1, False, ht, cs)
        B = q_idx.numel()
        out = ext.tt_forward(1000, 1, B, D, p, q, r, L, n_tt, col, rowidx, tbl, cores)
        if B - n_tt > 0:
            ext.cache_forward(B, B - n_tt, loc[n_tt:], rowidx[n_tt:], cw, out)
        outs[name] = (out, n_tt)
    assert _rel(outs["ours"][0], outs["ref"][0]) < TOL


def test_efficient_tt_matches_reference_kernels(ttg_lib):
    """Efficient_TT: Eff_TT_forward and Fused_Extra_Eff_TT_backward of the
    reference's own extension (efficient_tt_cuda.cu:243-377, :1011-1247) against the drop-in, on
    the products shape (the only BASELINE config its float index math is exact for)."""
    eff_ref = ref_ext.load_efficient()
    if eff_ref is None:
        pytest.skip("oracle/_ref Efficient_TT extension not built")
    import effi_tt_embeddings as eff
    p, q, r, n_emb = SHAPES["products"]
    D, batch = 100, 4096
    g = torch.Generator().manual_seed(9)
    cores0 = [(torch.rand(p[t], r[t] * q[t] * r[t + 1], generator=g) * 0.3).to(DEV) for t in range(3)]
    rng = np.random.default_rng(7)
    idx_np = rng.integers(0, 120000, size=batch).astype(np.int64)     # clustered: shared prefixes
    idx_np[:64] = idx_np[64:128]                                      # duplicates in the batch
    idx = torch.from_numpy(idx_np).to(DEV)
    tp = torch.tensor(p).to(DEV)
    tq = torch.tensor(q).to(DEV)
    tr = torch.tensor(r).to(DEV)
    eff_ref.init_cuda(0, q, r, batch, D)
    eff.init_cuda(0, q, r, batch, D)
    c_ref = [c.clone() for c in cores0]
    c_our = [c.clone() for c in cores0]
    want = eff_ref.Eff_TT_forward(batch, n_emb, D, idx, p, q, r, tp, tq, tr, c_ref)
    got = eff.Eff_TT_forward(batch, n_emb, D, idx, p, q, r, tp, tq, tr, c_our)
    torch.cuda.synchronize()
    assert got.shape == want.shape == (batch, D)
    assert _rel(got, want) < TOL
    dO = (torch.rand(batch, D, generator=g) * 0.1).to(DEV)
    uniq, inv = idx.unique(sorted=True, return_inverse=True)
    eff_ref.Fused_Extra_Eff_TT_backward(batch, n_emb, D, 0.1, idx, p, q, r, tp, tq, tr, dO.clone(),
                                        c_ref, uniq, inv)
    eff.Fused_Extra_Eff_TT_backward(batch, n_emb, D, 0.1, idx, p, q, r, tp, tq, tr, dO.clone(),
                                    c_our, uniq, inv)
    torch.cuda.synchronize()
    for t in range(3):
        assert not torch.equal(c_our[t], cores0[t])
    # compare the UPDATE (cores are O(0.3), the step is much smaller): ours against the fp64 oracle's gradient at
    # 1e-5 of the update's largest element, the reference's kernels (float atomics) within their stated 1e-4
    from oracle import oracle as orc
    truth = orc.tt_backward_dense(p, q, r, [c.cpu().numpy()[None] for c in cores0], idx_np,
                                  np.arange(batch, dtype=np.int64), dO.cpu().numpy()[None])
    for t in range(3):
        du, dr = (c_our[t] - cores0[t]).double(), (c_ref[t] - cores0[t]).double()
        dt = torch.from_numpy(-0.1 * truth[t][0].astype(np.float64)).to(DEV)
        scale = float(dt.abs().max())
        # cores of magnitude 0.3 minus cores of magnitude 0.3: the subtraction itself rounds at 0.3 * 6e-8
        floor = 0.3 * 1.2e-7 / scale
        assert float((du - dt).abs().max()) / scale < TOL + floor, "core %d: ours vs fp64 oracle" % t
        assert float((dr - dt).abs().max()) / scale < 1e-4, "core %d: reference kernels vs fp64 oracle" % t
