"""GPU parity tests: the CUDA path (through the C ABI) against the oracle, the committed golden
vectors of the reference, and size-independent properties at BASELINE.json's full sizes.

Tolerances: values 1e-5 relative (north star), index / hash / partition results bit-exact.
"""
import ctypes as C

import numpy as np
import pytest
import torch

from helpers import case_cores, collision_free_keys, random_block, rel_err
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

TOL = 1e-5
DEV = "cuda:0"

CASES = ["cora_r16", "cora_r16_bags", "products_small", "products_small_bags", "papers_small_r32",
         "two_cores", "four_cores"]


def _t(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    return t if dtype is None else t.to(dtype)


def _fwd(te, c, cores, indices, rowidx, B, num_tables=1, tableidx=None):
    p, q, r = list(c["p"]), list(c["q"]), list(c["ranks"])
    tb = torch.zeros_like(indices) if tableidx is None else tableidx
    return te.tt_forward(1000, num_tables, B, int(np.prod(q)), p, q, r, None, indices.numel(),
                         indices, rowidx, tb, cores)


@pytest.fixture(params=[0, 16, 1], ids=["mma", "ffma", "generic"])
def te(request, ttg_lib):
    """default = tensor-core kernels (3xTF32) where the shape has them, FLAG_FFMA = the fp32 FFMA
    sorted kernels, FLAG_FORCE_GENERIC = the shape-generic kernels"""
    import tt_embeddings
    tt_embeddings.EXTRA_FLAGS = request.param
    yield tt_embeddings
    tt_embeddings.EXTRA_FLAGS = 0


@pytest.mark.parametrize("name", CASES)
def test_forward_golden(te, golden, name):
    c = golden[name]
    cores = [_t(x) for x in case_cores(c)]
    B = c["offsets"].size - 1
    out = _fwd(te, c, cores, _t(c["indices"]), _t(c["rowidx"]), B)
    assert out.shape == (1, B, int(np.prod(c["q"])))
    assert rel_err(out[0].cpu().numpy(), c["out64"]) < TOL
    assert rel_err(out[0].cpu().numpy(), c["out"]) < TOL


@pytest.mark.parametrize("name", CASES)
def test_dense_backward_golden(te, golden, name):
    c = golden[name]
    p, q, r = list(c["p"]), list(c["q"]), list(c["ranks"])
    cores = [_t(x) for x in case_cores(c)]
    idx, row = _t(c["indices"]), _t(c["rowidx"])
    dO = _t(c["d_output"])[None].contiguous()
    d = te.tt_dense_backward(1000, int(np.prod(q)), p, q, r, None, idx.numel(), idx, row,
                             torch.zeros_like(idx), dO, cores)
    for t, g in enumerate(d):
        assert rel_err(g.cpu().numpy(), c["d_core%d" % t]) < TOL, "core %d" % t


@pytest.mark.parametrize("name", ["products_small_bags", "cora_r16", "four_cores"])
def test_fused_sgd_and_adagrad(te, golden, name):
    c = golden[name]
    p, q, r = list(c["p"]), list(c["q"]), list(c["ranks"])
    T = len(p)
    cols = [r[t] * q[t] * r[t + 1] for t in range(T)]
    idx, row = _t(c["indices"]), _t(c["rowidx"])
    dO = _t(c["d_output"])[None].contiguous()
    grads = [c["d_core%d" % t] for t in range(T)]
    # SGD: core - lr * grad on EVERY row (the reference skips tail rows)
    cores = [_t(x) for x in case_cores(c)]
    te.tt_sgd_backward(1000, int(np.prod(q)), 0.1, p, q, r, None, idx.numel(), idx, row,
                       torch.zeros_like(idx), dO, cores)
    want = [x.copy() for x in case_cores(c)]
    orc.apply_optimizer(p, cols, "sgd", 0.1, 0.0, want, None, grads)
    for t in range(T):
        assert rel_err(cores[t].cpu().numpy(), want[t]) < TOL
    # ... and on the rows the reference does update the two agree by construction
    lim = orc.reference_sgd_rows_updated(p, cols)
    want_ref = [x.copy() for x in case_cores(c)]
    orc.apply_optimizer(p, cols, "sgd", 0.1, 0.0, want_ref, None, grads, rows_limit=lim)
    for t in range(T):
        assert rel_err(cores[t].cpu().numpy()[:, :lim[t]], want_ref[t][:, :lim[t]]) < TOL
    # Adagrad, two steps so the state matters
    cores = [_t(x) for x in case_cores(c)]
    state = [torch.zeros_like(x) for x in cores]
    want = [x.copy() for x in case_cores(c)]
    wstate = [np.zeros_like(x) for x in want]
    for _ in range(2):
        te.tt_adagrad_backward(1000, int(np.prod(q)), 0.05, 1e-10, p, q, r, None, idx.numel(), idx,
                               row, torch.zeros_like(idx), dO, state, cores)
        g = orc.tt_backward_dense(p, q, r, want, c["indices"], c["rowidx"], c["d_output"][None])
        orc.apply_optimizer(p, cols, "adagrad", 0.05, 1e-10, want, wstate, g)
    # state = sum of g^2: 1e-5.  The Adagrad step lr * g / (sqrt(state) + 1e-10) is lr * sign(g) on the first
    # step whatever |g| is: where the true gradient is at rounding level its SIGN is not determined at fp32,
    # so the cores are held to 1e-5 where |g| >= 1e-3 max|g| and to one step (2 lr) elsewhere.
    g0 = orc.tt_backward_dense(p, q, r, case_cores(c), c["indices"], c["rowidx"], c["d_output"][None])
    for t in range(T):
        assert rel_err(state[t].cpu().numpy(), wstate[t]) < TOL
        well = np.abs(g0[t]) >= 1e-3 * np.abs(g0[t]).max()
        diff = np.abs(cores[t].cpu().numpy().astype(np.float64) - want[t])
        assert diff[well].max() < TOL * np.abs(want[t]).max()
        assert diff.max() <= 2 * 2 * 0.05 + 1e-6


def test_empty_and_degenerate_inputs(te, golden):
    c = golden["products_small"]
    p, q, r = list(c["p"]), list(c["q"]), list(c["ranks"])
    cores = [_t(x) for x in case_cores(c)]
    e = torch.empty(0, dtype=torch.int64, device=DEV)
    out = te.tt_forward(1000, 1, 7, 100, p, q, r, None, 0, e, e, e, cores)
    assert out.shape == (1, 7, 100) and float(out.abs().max()) == 0.0
    d = te.tt_dense_backward(1000, 100, p, q, r, None, 0, e, e, e,
                             torch.ones(1, 7, 100, device=DEV), cores)
    assert all(float(g.abs().max()) == 0.0 for g in d)
    # nnz smaller than the arrays: only the first nnz entries are used
    idx = _t(c["indices"])
    row = _t(c["rowidx"])
    out = te.tt_forward(1000, 1, 150, 100, p, q, r, None, 10, idx, row, torch.zeros_like(idx), cores)
    assert rel_err(out[0, :10].cpu().numpy(), c["out64"][:10]) < TOL
    assert float(out[0, 10:].abs().max()) == 0.0
    # a single row, and every index identical (one group, one bag)
    one = torch.tensor([int(np.prod(p)) - 1], device=DEV)
    out = te.tt_forward(1000, 1, 1, 100, p, q, r, None, 1, one, torch.zeros_like(one),
                        torch.zeros_like(one), cores)
    ref = orc.tt_forward(p, q, r, case_cores(c), one.cpu().numpy(), np.zeros(1, np.int64), 1)
    assert rel_err(out.cpu().numpy(), ref) < TOL
    same = torch.full((300,), 17, device=DEV, dtype=torch.int64)
    out = te.tt_forward(1000, 1, 1, 100, p, q, r, None, 300, same, torch.zeros_like(same),
                        torch.zeros_like(same), cores)
    assert rel_err(out.cpu().numpy(), 300.0 * orc.tt_forward(p, q, r, case_cores(c), [17], [0], 1)) < TOL


def test_multiple_tables(te):
    rng = np.random.default_rng(5)
    p, q, r = [6, 7, 5], [4, 5, 5], [1, 16, 16, 1]
    nt, B = 3, 40
    cores_np = [rng.normal(size=(nt, p[t], r[t] * q[t] * r[t + 1])).astype(np.float32) * 0.3
                for t in range(3)]
    lengths = rng.integers(0, 4, size=nt * B)
    offsets = np.concatenate([[0], np.cumsum(lengths)]).astype(np.int64)
    nnz = int(offsets[-1])
    idx = rng.integers(0, 6 * 7 * 5, size=nnz).astype(np.int64)
    _, rowidx, tableidx, _, _ = orc.preprocess_indices(idx, offsets, nt, True, None, None)
    want = orc.tt_forward(p, q, r, cores_np, idx, rowidx, B, tableidx, nt)
    cores = [_t(x) for x in cores_np]
    got = te.tt_forward(1000, nt, B, 100, p, q, r, None, nnz, _t(idx), _t(rowidx), _t(tableidx), cores)
    assert rel_err(got.cpu().numpy(), want) < TOL
    dO = rng.random(size=(nt, B, 100)).astype(np.float32) * 0.1
    wd = orc.tt_backward_dense(p, q, r, cores_np, idx, rowidx, dO, tableidx, nt)
    gd = te.tt_dense_backward(1000, 100, p, q, r, None, nnz, _t(idx), _t(rowidx), _t(tableidx),
                              _t(dO), cores)
    for t in range(3):
        assert rel_err(gd[t].cpu().numpy(), wd[t]) < TOL


@pytest.mark.parametrize("name", ["products_small_bags", "cora_r16_bags"])
def test_rowidx_in_any_order(te, golden, name):
    """The reference accumulates output[rowidx[n]] for any rowidx order
    (FBTT/tt_embeddings_cuda.cu:1027-1078); shuffling (indices, rowidx) together changes nothing."""
    c = golden[name]
    cores = [_t(x) for x in case_cores(c)]
    B = c["offsets"].size - 1
    perm = np.random.default_rng(4).permutation(c["indices"].size)
    out = _fwd(te, c, cores, _t(c["indices"][perm]), _t(c["rowidx"][perm]), B)
    assert rel_err(out[0].cpu().numpy(), c["out64"]) < TOL


def test_out_of_range_indices_contribute_nothing(te, golden):
    c = golden["products_small"]
    p, q, r = list(c["p"]), list(c["q"]), list(c["ranks"])
    cores = [_t(x) for x in case_cores(c)]
    idx = torch.tensor([3, -1, 10 ** 9, 5], device=DEV)
    row = torch.arange(4, device=DEV)
    out = te.tt_forward(1000, 1, 4, 100, p, q, r, None, 4, idx, row, torch.zeros_like(idx), cores)
    ref = orc.tt_forward(p, q, r, case_cores(c), [3, 5], [0, 3], 4)
    assert rel_err(out.cpu().numpy(), ref) < TOL


# --------------------------------------------------------------------------------------------
# medium sizes against the oracle, full sizes through properties
# --------------------------------------------------------------------------------------------
SHAPES = {
    "products": ([125, 140, 140], [4, 5, 5], [1, 16, 16, 1], 2449029),
    "arxiv": ([55, 55, 56], [4, 4, 8], [1, 16, 16, 1], 169343),
    "papers": ([481, 481, 481], [4, 4, 8], [1, 32, 32, 1], 111059956),
    "cora": ([14, 14, 14], [4, 4, 8], [1, 16, 16, 1], 2708),
    "arxiv_q844": ([55, 55, 56], [8, 4, 4], [1, 16, 16, 1], 169343),     # run_script.sh:299,316
    "products_q545": ([125, 140, 140], [5, 4, 5], [1, 16, 16, 1], 2449029),  # run_script.sh:353
    # run_script.sh:247-264: the tt-ranks sweep at products shape (--q-shapes "5,5,4"; ranks 64 and up stay on
    # the any-shape kernels)
    "products_q554": ([125, 140, 140], [5, 5, 4], [1, 16, 16, 1], 2449029),
    "products_q554_r8": ([125, 140, 140], [5, 5, 4], [1, 8, 8, 1], 2449029),
    "products_q554_r32": ([125, 140, 140], [5, 5, 4], [1, 32, 32, 1], 2449029),
}


def _random_cores(p, q, r, n_emb, seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(1, p[t], r[t] * q[t] * r[t + 1], generator=g) * (0.5 / np.sqrt(r[t]))
            for t in range(3)]


@pytest.mark.parametrize("shape", ["products", "arxiv", "papers", "cora", "arxiv_q844", "products_q545",
                                   "products_q554", "products_q554_r8", "products_q554_r32"])
def test_medium_batch_against_oracle(ttg_lib, shape):
    import _ttg
    import tt_embeddings as te
    # the shape runs on the plan-based kernels, not on the any-shape fallback (TTG_ENOTSUP = -4 there)
    assert _ttg.lib().ttg_tt_plan(C.byref(_ttg.make_shape(*SHAPES[shape][:3])), 1, 0, None, None, None, None, 0, 0,
                                  None) == 0
    p, q, r, n_emb = SHAPES[shape]
    D = int(np.prod(q))
    cores_cpu = _random_cores(p, q, r, n_emb, 11)
    cores = [c.to(DEV) for c in cores_cpu]
    rng = np.random.default_rng(3)
    nnz = 6000
    idx = rng.integers(0, n_emb, size=nnz).astype(np.int64)
    idx[: nnz // 3] = np.sort(rng.integers(0, min(n_emb, 4000), size=nnz // 3))  # clustered part
    idx[5] = idx[6] = idx[7]
    row = np.arange(nnz, dtype=np.int64)
    dO = (rng.random(size=(1, nnz, D)).astype(np.float32) * 0.1)
    cn = [c.numpy() for c in cores_cpu]
    want = orc.tt_forward(p, q, r, cn, idx, row, nnz)
    got = te.tt_forward(1000, 1, nnz, D, p, q, r, None, nnz, _t(idx), _t(row),
                        torch.zeros(nnz, dtype=torch.int64, device=DEV), cores)
    assert rel_err(got.cpu().numpy(), want) < TOL
    wd = orc.tt_backward_dense(p, q, r, cn, idx, row, dO)
    gd = te.tt_dense_backward(1000, D, p, q, r, None, nnz, _t(idx), _t(row),
                              torch.zeros(nnz, dtype=torch.int64, device=DEV), _t(dO), cores)
    for t in range(3):
        assert rel_err(gd[t].cpu().numpy(), wd[t]) < TOL, "core %d" % t


@pytest.mark.parametrize("base_flags", [0, 16], ids=["mma", "ffma"])
def test_full_size_products_properties(ttg_lib, base_flags):
    """BASELINE config 2 size (262,144 distinct ids of 2,449,029): sorted kernels == generic
    kernels, permutation equivariance, linearity of the gradient in d_output."""
    import _ttg
    import tt_embeddings as te
    te.EXTRA_FLAGS = base_flags
    p, q, r, n_emb = SHAPES["products"]
    D = 100
    cores = [c.to(DEV) for c in _random_cores(p, q, r, n_emb, 12)]
    g = torch.Generator(device="cpu").manual_seed(0)
    nnz = 262144
    idx = torch.randperm(n_emb, generator=g)[:nnz].to(DEV)
    row = torch.arange(nnz, device=DEV)
    tb = torch.zeros_like(idx)
    out = te.tt_forward(1000, 1, nnz, D, p, q, r, None, nnz, idx, row, tb, cores)
    te.EXTRA_FLAGS = _ttg.FLAG_FORCE_GENERIC
    try:
        out_g = te.tt_forward(1000, 1, nnz, D, p, q, r, None, nnz, idx, row, tb, cores)
    finally:
        te.EXTRA_FLAGS = base_flags
    assert float((out - out_g).abs().max() / out_g.abs().max()) < TOL
    # permutation equivariance: looking up a permuted index list permutes the rows
    perm = torch.randperm(nnz, generator=g).to(DEV)
    out_p = te.tt_forward(1000, 1, nnz, D, p, q, r, None, nnz, idx[perm].contiguous(), row, tb, cores)
    assert torch.equal(out_p[0], out[0][perm])
    # sorted ids give the same rows
    sidx, order = torch.sort(idx)
    out_s = te.tt_forward(1000, 1, nnz, D, p, q, r, None, nnz, sidx, row, tb, cores)
    assert torch.equal(out_s[0], out[0][order])
    # gradient: sorted path vs generic path, and linearity in d_output
    dO = torch.rand(1, nnz, D, generator=g).to(DEV) * 0.1
    gs = te.tt_dense_backward(1000, D, p, q, r, None, nnz, idx, row, tb, dO, cores)
    te.EXTRA_FLAGS = _ttg.FLAG_FORCE_GENERIC
    try:
        gg = te.tt_dense_backward(1000, D, p, q, r, None, nnz, idx, row, tb, dO, cores)
    finally:
        te.EXTRA_FLAGS = base_flags
    # against the fp64 oracle at the full batch: 1e-5; the generic kernels scatter with float atomics (as the
    # reference does) and are the ones that need the wider, stated bound
    orc.use_all_host_threads()
    truth = orc.tt_backward_dense(p, q, r, [c.cpu().numpy() for c in cores], idx.cpu().numpy(),
                                  row.cpu().numpy(), dO.cpu().numpy())
    for t, (a, b) in enumerate(zip(gs, gg)):
        assert rel_err(a.cpu().numpy(), truth[t]) < TOL, "core %d vs fp64 oracle" % t
        assert float((a - b).abs().max() / b.abs().max()) < 5e-5   # generic path: float atomics
    g2 = te.tt_dense_backward(1000, D, p, q, r, None, nnz, idx, row, tb, (2.0 * dO).contiguous(), cores)
    for a, b in zip(gs, g2):
        assert float((2.0 * a - b).abs().max() / b.abs().max()) < 1e-6
    # determinism of the sorted path (no atomics in the gradient reductions except shared-memory
    # accumulation order inside a CTA)
    gs2 = te.tt_dense_backward(1000, D, p, q, r, None, nnz, idx, row, tb, dO, cores)
    for a, b in zip(gs, gs2):   # bucket plan: order inside a group is arbitrary -> fp32 reorder
        assert float((a - b).abs().max() / b.abs().max()) < 1e-6
    # TTG_FLAG_DETERMINISTIC: radix-sorted plan, fixed summation order for core0 / core1
    te.EXTRA_FLAGS = base_flags | _ttg.FLAG_DETERMINISTIC
    try:
        gd1 = te.tt_dense_backward(1000, D, p, q, r, None, nnz, idx, row, tb, dO, cores)
        gd2 = te.tt_dense_backward(1000, D, p, q, r, None, nnz, idx, row, tb, dO, cores)
    finally:
        te.EXTRA_FLAGS = base_flags
    assert torch.equal(gd1[0], gd2[0]) and torch.equal(gd1[1], gd2[1])
    assert float((gd1[2] - gd2[2]).abs().max() / gd1[2].abs().max()) < 1e-6
    for a, b in zip(gs, gd1):
        assert float((a - b).abs().max() / b.abs().max()) < 1e-6
    te.EXTRA_FLAGS = 0


def test_tf32_mode_stated_bound(ttg_lib):
    """TTG_FLAG_TF32: single-pass TF32 tensor-core kernels.  Stated bound: 3e-3 of the tensor's
    max (TF32 has a 10-bit mantissa; the default 3xTF32 mode is held to 1e-5 everywhere else)."""
    import _ttg
    import tt_embeddings as te
    p, q, r, n_emb = SHAPES["products"]
    D = 100
    cores_cpu = _random_cores(p, q, r, n_emb, 21)
    cores = [c.to(DEV) for c in cores_cpu]
    rng = np.random.default_rng(5)
    nnz = 40000
    idx = rng.integers(0, 60000, size=nnz).astype(np.int64)     # ~430 groups, dense in groups
    row = np.arange(nnz, dtype=np.int64)
    dO = (rng.random(size=(1, nnz, D)).astype(np.float32) * 0.1)
    cn = [c.numpy() for c in cores_cpu]
    want = orc.tt_forward(p, q, r, cn, idx, row, nnz)
    wd = orc.tt_backward_dense(p, q, r, cn, idx, row, dO)
    tb = torch.zeros(nnz, dtype=torch.int64, device=DEV)
    for flags, tol in ((_ttg.FLAG_TF32, 3e-3), (0, TOL)):
        te.EXTRA_FLAGS = flags
        try:
            got = te.tt_forward(1000, 1, nnz, D, p, q, r, None, nnz, _t(idx), _t(row), tb, cores)
            gd = te.tt_dense_backward(1000, D, p, q, r, None, nnz, _t(idx), _t(row), tb, _t(dO), cores)
        finally:
            te.EXTRA_FLAGS = 0
        assert rel_err(got.cpu().numpy(), want) < tol
        for t in range(3):
            assert rel_err(gd[t].cpu().numpy(), wd[t]) < tol, "core %d" % t


def test_full_table_arange_arxiv(ttg_lib):
    """Config 4 call pattern: every row of the table, sorted (gcn_gat_partition.py:93-96)."""
    import tt_embeddings as te
    from FBTT.tt_embeddings_ops import tt_matrix_to_full
    p, q, r, n_emb = SHAPES["arxiv"]
    cores = [c.to(DEV) for c in _random_cores(p, q, r, n_emb, 13)]
    idx = torch.arange(n_emb, device=DEV)
    out = te.tt_forward(1000, 1, n_emb, 128, p, q, r, None, n_emb, idx, idx, torch.zeros_like(idx), cores)
    W = tt_matrix_to_full(p, q, r, cores, [1, 0, 2, 3])[:n_emb]
    assert float((out[0] - W).abs().max() / W.abs().max()) < TOL


# --------------------------------------------------------------------------------------------
# index path: bit-exact against the oracle
# --------------------------------------------------------------------------------------------
def test_cache_index_path_bit_exact(ttg_lib, golden):
    import tt_embeddings as te
    c = golden["products_small"]
    p, q, r = list(c["p"]), list(c["q"]), list(c["ranks"])
    n_emb, D = int(np.prod(p)), 100
    rng = np.random.default_rng(9)
    size, cache_size = 4001, 12
    keys = collision_free_keys(orc, size, 60, rng, 0, n_emb)
    counts = rng.integers(1, 9, size=keys.size)
    counts[:3] = [20, 20, 19]                       # ties among the most frequent
    stream = np.repeat(keys, counts)
    rng.shuffle(stream)
    h_k = np.full(size, -1, np.int64)
    h_f = np.zeros(size, np.int64)
    h_s = np.full(size, -1, np.int32)
    d_k, d_f, d_s = _t(h_k), _t(h_f), _t(h_s)
    orc.update_cache_state(stream, h_k, h_f)
    te.update_cache_state(_t(stream), d_k, d_f)
    assert np.array_equal(d_k.cpu().numpy(), h_k) and np.array_equal(d_f.cpu().numpy(), h_f)
    # populate
    cores = [_t(x) for x in case_cores(c)]
    cw = torch.zeros(cache_size, D, device=DEV)
    te.cache_populate(n_emb, p, q, r, cores, None, d_k, d_f, d_s, cw)
    sorted_keys = orc.cache_populate_index(cache_size, h_k, h_f, h_s)
    assert np.array_equal(d_k.cpu().numpy(), h_k)
    assert np.array_equal(d_f.cpu().numpy(), h_f)
    assert np.array_equal(d_s.cpu().numpy(), h_s)
    rows = orc.tt_forward(p, q, r, case_cores(c), sorted_keys[:cache_size], np.arange(cache_size),
                          cache_size)[0]
    assert rel_err(cw.cpu().numpy(), rows) < TOL
    # lookup + partition, with bags
    lengths = rng.integers(0, 4, size=300)
    offsets = np.concatenate([[0], np.cumsum(lengths)]).astype(np.int64)
    nnz = int(offsets[-1])
    col = rng.choice(np.concatenate([keys, rng.integers(0, n_emb, size=40)]), size=nnz)
    want = orc.preprocess_indices(col, offsets, 1, False, h_k, h_s)
    got = te.preprocess_indices_sync(_t(col), _t(offsets), 1, False, d_k, d_s)
    assert got[3] == want[3]
    for a, b in zip((got[0], got[1], got[2], got[4]), (want[0], want[1], want[2], want[4])):
        assert np.array_equal(a.cpu().numpy(), b)
    assert 0 < got[3] < nnz
    # warm-up path returns the inputs
    got_w = te.preprocess_indices_sync(_t(col), _t(offsets), 1, True, d_k, d_s)
    assert got_w[3] == nnz and got_w[4] is None
    assert np.array_equal(got_w[1].cpu().numpy(),
                          np.repeat(np.arange(300, dtype=np.int64), lengths))
    # cached rows forward / backward
    n_tt = got[3]
    loc, row = got[4][n_tt:], got[1][n_tt:]
    out_h = np.random.default_rng(1).normal(size=(300, D)).astype(np.float32)
    out_d = _t(out_h)
    te.cache_forward(300, nnz - n_tt, loc, row, cw, out_d)
    orc.cache_forward(loc.cpu().numpy(), row.cpu().numpy(), cw.cpu().numpy(), out_h)
    assert rel_err(out_d.cpu().numpy(), out_h) < 1e-6
    grad = np.random.default_rng(2).normal(size=(1, 300, D)).astype(np.float32)
    w_h = cw.cpu().numpy().copy()
    orc.cache_backward(grad.reshape(300, D), loc.cpu().numpy(), row.cpu().numpy(), 0.1, 0, w_h)
    w_d = cw.clone()
    te.cache_backward_sgd(nnz - n_tt, _t(grad), loc, row, 0.1, w_d)
    assert rel_err(w_d.cpu().numpy(), w_h) < TOL
    g_h = np.zeros((cache_size, D), np.float32)
    orc.cache_backward(grad.reshape(300, D), loc.cpu().numpy(), row.cpu().numpy(), 0.0, 1, g_h)
    g_d = te.cache_backward_dense(nnz - n_tt, _t(grad), loc, row, 0.1, cw)
    assert rel_err(g_d.cpu().numpy(), g_h) < TOL
    # row-wise adagrad: use distinct cache rows so the update order cannot matter
    uniq_loc, first = np.unique(loc.cpu().numpy(), return_index=True)
    l1 = _t(loc.cpu().numpy()[first])
    r1 = _t(row.cpu().numpy()[first])
    order = torch.argsort(r1, descending=True)
    l1, r1 = l1[order].contiguous(), r1[order].contiguous()
    keep = np.concatenate([[True], np.diff(r1.cpu().numpy()) != 0])   # one index per row segment
    l1, r1 = l1[_t(keep)].contiguous(), r1[_t(keep)].contiguous()
    st_h = np.zeros(cache_size, np.float32)
    w_h = cw.cpu().numpy().copy()
    orc.cache_backward_rowwise_adagrad(grad.reshape(300, D), l1.cpu().numpy(), r1.cpu().numpy(), 0.1,
                                       1e-8, st_h, w_h)
    st_d = torch.zeros(cache_size, device=DEV)
    w_d = cw.clone()
    te.cache_backward_rowwise_adagrad_approx(l1.numel(), _t(grad), l1, r1, 0.1, 1e-8, st_d, w_d)
    assert rel_err(st_d.cpu().numpy(), st_h) < TOL and rel_err(w_d.cpu().numpy(), w_h) < TOL


def test_hash_insert_under_collisions_as_sets(ttg_lib):
    """Load factor 1 (hashtbl_size = num_nodes, every node touched): which colliding key wins a
    slot is thread-order dependent in the reference too; compare invariants instead."""
    import tt_embeddings as te
    size = 20000
    idx = torch.randperm(size, device=DEV)
    d_k = torch.full((size,), -1, dtype=torch.int64, device=DEV)
    d_f = torch.zeros(size, dtype=torch.int64, device=DEV)
    te.update_cache_state(idx, d_k, d_f)
    te.update_cache_state(idx, d_k, d_f)
    k = d_k.cpu().numpy()
    f = d_f.cpu().numpy()
    present = k[k != -1]
    assert present.size == np.unique(present).size          # no key stored twice
    assert set(f[k != -1]) == {2} and (f[k == -1] == 0).all()
    for slot in np.nonzero(k != -1)[0][:2000]:               # every key sits within 3 probes
        h = orc.hash32(int(k[slot]), size)
        assert (slot - h) % size < 3
    assert 0.70 < present.size / size < 0.85                 # ~22 % cannot be placed (SURVEY A)


# --------------------------------------------------------------------------------------------
# aggregation
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("F", [100, 256, 48])
@pytest.mark.parametrize("mean", [True, False])
def test_spmm(ttg_lib, F, mean):
    import gnn_ops
    rng = np.random.default_rng(F)
    num_src, num_dst = 3000, 700
    indptr, indices = random_block(rng, num_src, num_dst, 17)
    x = rng.normal(size=(num_src, F)).astype(np.float32)
    blk = gnn_ops.Block(_t(indptr), _t(indices), num_src, num_dst)
    xt = _t(x).requires_grad_(True)
    out = gnn_ops.aggregate(blk, xt, mean)
    assert rel_err(out.detach().cpu().numpy(), orc.spmm_csr_fwd(indptr, indices, x, mean)) < TOL
    dout = rng.normal(size=(num_dst, F)).astype(np.float32)
    out.backward(_t(dout))
    assert rel_err(xt.grad.cpu().numpy(), orc.spmm_csr_bwd(indptr, indices, dout, num_src, mean)) < TOL


# --------------------------------------------------------------------------------------------
# tensor-core path: adversarial index distributions (chunk ownership, tails, invalid keys)
# --------------------------------------------------------------------------------------------
def _mma_case(te, shape, idx, B=None, row=None, dO_seed=1, tol=TOL):
    p, q, r, n_emb = SHAPES[shape]
    D = int(np.prod(q))
    cores_cpu = _random_cores(p, q, r, n_emb, 31)
    cores = [c.to(DEV) for c in cores_cpu]
    cn = [c.numpy() for c in cores_cpu]
    nnz = idx.size
    row = np.arange(nnz, dtype=np.int64) if row is None else row
    B = nnz if B is None else B
    valid = (idx >= 0) & (idx < int(np.prod(p)))
    want = orc.tt_forward(p, q, r, cn, idx[valid], row[valid], B)
    tb = torch.zeros(nnz, dtype=torch.int64, device=DEV)
    got = te.tt_forward(1000, 1, B, D, p, q, r, None, nnz, _t(idx), _t(row), tb, cores)
    assert rel_err(got.cpu().numpy(), want) < tol
    dO = (np.random.default_rng(dO_seed).random(size=(1, B, D)).astype(np.float32) * 0.1)
    wd = orc.tt_backward_dense(p, q, r, cn, idx[valid], row[valid], dO)
    gd = te.tt_dense_backward(1000, D, p, q, r, None, nnz, _t(idx), _t(row), tb, _t(dO), cores)
    for t in range(3):
        assert rel_err(gd[t].cpu().numpy(), wd[t]) < tol, "core %d" % t


@pytest.mark.parametrize("shape", ["products", "arxiv"])
def test_mma_one_giant_group_and_duplicates(ttg_lib, shape):
    """All rows in one (i0, i1) group -- a single warp owns the whole batch's group -- and a
    batch that is one index repeated; plus both mixed with a uniform background."""
    import tt_embeddings as te
    p, q, r, n_emb = SHAPES[shape]
    rng = np.random.default_rng(17)
    g0 = 37 * p[2]
    one_group = (g0 + rng.integers(0, p[2], size=9000)).astype(np.int64)
    _mma_case(te, shape, one_group)
    same = np.full(7000, g0 + 3, dtype=np.int64)
    _mma_case(te, shape, same)
    ngroups = p[0] * p[1]
    mixed = np.concatenate([one_group[:4000], same[:1000],
                            rng.integers(0, n_emb, size=max(ngroups, 6000))]).astype(np.int64)
    rng.shuffle(mixed)
    _mma_case(te, shape, mixed)


def test_mma_tiny_and_ragged_batches(ttg_lib):
    import tt_embeddings as te
    p, q, r, n_emb = SHAPES["cora"]       # 196 groups: the table strategy starts at nnz = 196
    rng = np.random.default_rng(23)
    for nnz in (196, 197, 211, 255, 513, 1000):
        _mma_case(te, "cora", rng.integers(0, n_emb, size=nnz).astype(np.int64))


def test_mma_invalid_indices_and_bags(ttg_lib):
    """Out-of-range / negative ids sort behind every valid key and contribute nothing; bags with
    0..7 indices exercise the zero-fill and the accumulate path of the bulk-copy stores."""
    import tt_embeddings as te
    p, q, r, n_emb = SHAPES["arxiv"]
    rng = np.random.default_rng(29)
    lengths = rng.integers(0, 8, size=3000)
    nnz = int(lengths.sum())
    idx = rng.integers(0, n_emb, size=nnz).astype(np.int64)
    bad = rng.permutation(nnz)[:200]
    idx[bad[:100]] = -1 - rng.integers(0, 5, size=100)
    idx[bad[100:]] = int(np.prod(p)) + rng.integers(0, 1000, size=100)
    row = np.repeat(np.arange(lengths.size), lengths).astype(np.int64)
    _mma_case(te, "arxiv", idx, B=lengths.size, row=row)


def test_rank32_tensor_core_forward_small_against_oracle(ttg_lib):
    """ranks 32, 32 (papers100M recipe) on a table small enough that every group has rows: the
    group table + tensor-core forward (core2 fragments from global memory) against the oracle,
    then the FFMA backward on the plan the forward left behind."""
    import tt_embeddings as te
    p, q, r = [5, 6, 7], [4, 4, 8], [1, 32, 32, 1]
    n_emb, D = 5 * 6 * 7, 128
    cores_cpu = _random_cores(p, q, r, n_emb, 21)
    cores = [c.to(DEV) for c in cores_cpu]
    rng = np.random.default_rng(8)
    nnz = 700
    idx = rng.integers(0, n_emb, size=nnz).astype(np.int64)
    row = rng.permutation(nnz).astype(np.int64)
    tb = torch.zeros(nnz, dtype=torch.int64, device=DEV)
    cn = [c.numpy() for c in cores_cpu]
    want = orc.tt_forward(p, q, r, cn, idx, row, nnz)
    got = te.tt_forward(1000, 1, nnz, D, p, q, r, None, nnz, _t(idx), _t(row), tb, cores)
    assert rel_err(got.cpu().numpy(), want) < TOL
    te.EXTRA_FLAGS = 16
    try:
        got_ffma = te.tt_forward(1000, 1, nnz, D, p, q, r, None, nnz, _t(idx), _t(row), tb, cores)
    finally:
        te.EXTRA_FLAGS = 0
    assert rel_err(got_ffma.cpu().numpy(), want) < TOL
    dO = rng.random(size=(1, nnz, D)).astype(np.float32) * 0.1
    wd = orc.tt_backward_dense(p, q, r, cn, idx, row, dO)
    gd = te.tt_dense_backward(1000, D, p, q, r, None, nnz, _t(idx), _t(row), tb, _t(dO), cores)
    for t in range(3):
        assert rel_err(gd[t].cpu().numpy(), wd[t]) < TOL, "core %d" % t


def test_full_size_papers_forward_properties(ttg_lib):
    """BASELINE config 5 size (262,144 ids of 111,059,956, ids above 2^24 and duplicates included):
    tensor-core forward == FFMA forward == generic kernels; permutation equivariance."""
    import _ttg
    import tt_embeddings as te
    p, q, r, n_emb = SHAPES["papers"]
    D = 128
    cores = [c.to(DEV) for c in _random_cores(p, q, r, n_emb, 13)]
    g = torch.Generator(device="cpu").manual_seed(1)
    nnz = 262144
    idx = torch.randint(0, n_emb, (nnz,), generator=g)
    idx[:1000] = idx[1000:2000]                       # duplicates
    idx[2000:2200] = torch.arange(n_emb - 200, n_emb)  # the last rows of the table
    idx = idx.to(DEV)
    row = torch.arange(nnz, device=DEV)
    tb = torch.zeros_like(idx)
    out = te.tt_forward(1000, 1, nnz, D, p, q, r, None, nnz, idx, row, tb, cores)
    outs = {}
    for name, fl in (("ffma", 16), ("generic", _ttg.FLAG_FORCE_GENERIC)):
        te.EXTRA_FLAGS = fl
        try:
            outs[name] = te.tt_forward(1000, 1, nnz, D, p, q, r, None, nnz, idx, row, tb, cores)
        finally:
            te.EXTRA_FLAGS = 0
    for name, o in outs.items():
        assert float((out - o).abs().max() / o.abs().max()) < TOL, name
    perm = torch.randperm(nnz, generator=g).to(DEV)
    out_p = te.tt_forward(1000, 1, nnz, D, p, q, r, None, nnz, idx[perm].contiguous(), row, tb, cores)
    assert torch.equal(out_p[0], out[0][perm])
    # the backward (FFMA kernels at these ranks) on the plan the tensor-core forward left behind
    dO = torch.rand(1, nnz, D, generator=g).to(DEV) * 0.1
    out = te.tt_forward(1000, 1, nnz, D, p, q, r, None, nnz, idx, row, tb, cores)
    gs = te.tt_dense_backward(1000, D, p, q, r, None, nnz, idx, row, tb, dO, cores)
    te.EXTRA_FLAGS = _ttg.FLAG_FORCE_GENERIC
    try:
        gg = te.tt_dense_backward(1000, D, p, q, r, None, nnz, idx, row, tb, dO, cores)
    finally:
        te.EXTRA_FLAGS = 0
    for t in range(3):
        assert float((gs[t] - gg[t]).abs().max() / gg[t].abs().max()) < 5e-5, "core %d" % t


def test_back_to_back_steps_with_a_deep_launch_queue(ttg_lib):
    """The row / cores / finalize kernels start early (programmatic dependent launch) and stage
    their operands before the kernel in front of them has finished.  Thirty forward + fused-SGD
    steps are queued behind a busy GPU without any host synchronisation -- every kernel finds its
    predecessor still running -- and must leave the cores the generic kernels leave."""
    import _ttg
    import tt_embeddings as te
    p, q, r, n_emb = SHAPES["products"]
    D, nnz, steps = 100, 30000, 30
    g = torch.Generator(device="cpu").manual_seed(5)
    batches = [torch.randint(0, n_emb, (nnz,), generator=g).to(DEV) for _ in range(4)]
    dOs = [(torch.rand(1, nnz, D, generator=g) * 0.1).to(DEV) for _ in range(4)]
    row = torch.arange(nnz, device=DEV)
    tb = torch.zeros(nnz, dtype=torch.int64, device=DEV)
    start = [c.to(DEV) for c in _random_cores(p, q, r, n_emb, 14)]
    results = {}
    for name, fl in (("generic", _ttg.FLAG_FORCE_GENERIC), ("mma", 0), ("backward_only", 0)):
        cores = [c.clone() for c in start]
        te.EXTRA_FLAGS = fl
        try:
            torch.cuda.synchronize()
            torch.cuda._sleep(40_000_000)        # about 20 ms: the queue fills behind it
            for s in range(steps):
                k = s % 4
                if name != "backward_only":
                    te.tt_forward(1000, 1, nnz, D, p, q, r, None, nnz, batches[k], row, tb, cores)
                te.tt_sgd_backward(1000, D, 2e-4, p, q, r, None, nnz, batches[k], row, tb, dOs[k], cores)
            torch.cuda.synchronize()
        finally:
            te.EXTRA_FLAGS = 0
        results[name] = cores
    for name in ("mma", "backward_only"):
        for t in range(3):
            upd = (results["generic"][t] - start[t]).abs().max()
            err = (results[name][t] - results["generic"][t]).abs().max()
            assert float(err / upd) < 1e-4, "%s core %d: %.3g of the update" % (name, t, float(err / upd))


def test_backward_only_after_a_call_with_another_row_count(ttg_lib):
    """Regression: the backward row kernel once read the plan's row count with an invariant
    (ld.global.nc) load that ptxas hoisted above griddepcontrol.wait, i.e. before the scan kernel
    of the same call had written it -- the count of the PREVIOUS call was used.  Backward-only
    calls with alternating row counts, queued behind a busy GPU."""
    import _ttg
    import tt_embeddings as te
    p, q, r, n_emb = SHAPES["products"]
    D = 100
    g = torch.Generator(device="cpu").manual_seed(6)
    sizes = [30000, 18000, 41000, 20000]
    idx = [torch.randint(0, n_emb, (n,), generator=g).to(DEV) for n in sizes]
    dOs = [(torch.rand(1, n, D, generator=g) * 0.1).to(DEV) for n in sizes]
    cores = [c.to(DEV) for c in _random_cores(p, q, r, n_emb, 15)]
    outs = {}
    for name, fl in (("generic", _ttg.FLAG_FORCE_GENERIC), ("mma", 0)):
        te.EXTRA_FLAGS = fl
        try:
            torch.cuda.synchronize()
            torch.cuda._sleep(40_000_000)
            res = []
            for k, n in enumerate(sizes):
                row = torch.arange(n, device=DEV)
                res.append(te.tt_dense_backward(1000, D, p, q, r, None, n, idx[k], row,
                                                torch.zeros_like(row), dOs[k], cores))
            torch.cuda.synchronize()
        finally:
            te.EXTRA_FLAGS = 0
        outs[name] = res
    for k in range(len(sizes)):
        for t in range(3):
            a, b = outs["mma"][k][t], outs["generic"][k][t]
            assert float((a - b).abs().max() / b.abs().max()) < 5e-5, "call %d core %d" % (k, t)


def test_plan_is_rebuilt_after_a_raw_pointer_core_update(ttg_lib):
    """The index plan and group table of a batch are reused by the backward that follows its forward
    (TTG_FLAG_PLAN_VALID, keyed on tensor version counters).  dp.apply_optimizer / the peer exchange update the
    cores through raw pointers, which no version counter sees: they must drop the plan themselves, or the next
    backward on the same indices would contract with the group table of the OLD cores."""
    import dp
    import tt_embeddings as te
    p, q, r, n_emb = SHAPES["cora"]
    D = 128
    cores_cpu = _random_cores(p, q, r, n_emb, 31)
    cores = [c.to(DEV) for c in cores_cpu]
    rng = np.random.default_rng(9)
    nnz = 3000
    idx = rng.integers(0, n_emb, size=nnz).astype(np.int64)
    row = np.arange(nnz, dtype=np.int64)
    dO = (rng.random(size=(1, nnz, D)).astype(np.float32) * 0.1)
    ti, tr, tb, tdo = _t(idx), _t(row), torch.zeros(nnz, dtype=torch.int64, device=DEV), _t(dO)
    te.tt_forward(1000, 1, nnz, D, p, q, r, None, nnz, ti, tr, tb, cores)
    d1 = te.tt_dense_backward(1000, D, p, q, r, None, nnz, ti, tr, tb, tdo, cores)
    dp.apply_optimizer(p, q, r, cores, [g.clone() for g in d1], 0.5)      # cores change, versions do not
    d2 = te.tt_dense_backward(1000, D, p, q, r, None, nnz, ti, tr, tb, tdo, cores)
    want = orc.tt_backward_dense(p, q, r, [c.cpu().numpy() for c in cores], idx, row, dO)
    for t in range(3):
        assert rel_err(d2[t].cpu().numpy(), want[t]) < TOL, "core %d: stale group table" % t


@pytest.mark.parametrize("nnz", [50000, 400000], ids=["left_grouped", "right_grouped"])
def test_plan_built_ahead_on_another_stream(ttg_lib, nnz):
    """(at 400,000 rows per call the right-grouped kernels run and the plan is sorted by their transposed keys)
    ttg_tt_plan: the index plan of the next batch built on a side stream into the other plan slot while the
    current batch runs; the forward / backward that follow recognise it (same tensors) and give bit-identical
    results to the calls that plan for themselves.  A batch of another size is refused (the workspace layout
    depends on nnz) and then simply plans for itself."""
    import _ttg
    import tt_embeddings as te
    p, q, r, n_emb = SHAPES["products"]
    D = 100
    cores = [c.to(DEV) for c in _random_cores(p, q, r, n_emb, 41)]
    g = torch.Generator().manual_seed(3)
    batches = [torch.randint(0, n_emb, (nnz,), generator=g).to(DEV) for _ in range(3)]
    dOs = [(torch.rand(1, nnz, D, generator=g) * 0.1).to(DEV) for _ in range(3)]
    row = torch.arange(nnz, device=DEV)
    tb = torch.zeros_like(row)
    want = []
    for k in range(3):
        o = te.tt_forward(1000, 1, nnz, D, p, q, r, None, nnz, batches[k], row, tb, cores)
        d = te.tt_dense_backward(1000, D, p, q, r, None, nnz, batches[k], row, tb, dOs[k], cores)
        want.append((o.clone(), [x.clone() for x in d]))
    side = torch.cuda.Stream()
    cur = torch.cuda.current_stream()
    assert te.tt_plan(1, nnz, p, q, r, nnz, batches[0], row, tb, 0)
    for k in range(3):
        if k + 1 < 3:                       # plan of batch k + 1 beside batch k
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                assert te.tt_plan(1, nnz, p, q, r, nnz, batches[k + 1], row, tb, (k + 1) & 1)
        assert _ttg.workspace.ready_slot(torch.device(DEV), _ttg.index_key_of(
            te._plan_tag(), batches[k], row, nnz, nnz, _ttg.make_shape(p, q, r, 1).key)) == (k & 1)
        o = te.tt_forward(1000, 1, nnz, D, p, q, r, None, nnz, batches[k], row, tb, cores)
        d = te.tt_dense_backward(1000, D, p, q, r, None, nnz, batches[k], row, tb, dOs[k], cores)
        cur.wait_stream(side)
        assert torch.equal(o, want[k][0])
        for a, b in zip(d, want[k][1]):
            # the bucket plan orders the rows of a group by atomics: fp32 sums reorder between two plans
            assert float((a - b).abs().max() / b.abs().max()) < 1e-6
    # another batch size: not prepared (the layout would differ), the forward plans for itself
    small = batches[0][:30000].contiguous()
    assert not te.tt_plan(1, 30000, p, q, r, 30000, small, row[:30000].contiguous(), tb[:30000].contiguous(), 1)
    o = te.tt_forward(1000, 1, 30000, D, p, q, r, None, 30000, small, row[:30000].contiguous(),
                      tb[:30000].contiguous(), cores)
    if nnz < 274400:     # 14 rows per (i1, i2) group: where the right-grouped kernels take over
        assert torch.equal(o[0], want[0][0][0, :30000])
    else:     # 30,000 rows run on the left-grouped kernels, the 400,000 of `want` on the right-grouped ones
        assert float((o[0] - want[0][0][0, :30000]).abs().max() / want[0][0].abs().max()) < TOL


def test_module_prepare_matches_plain_forward(ttg_lib):
    from FBTT.tt_embeddings_ops import OptimType, TTEmbeddingBag
    torch.manual_seed(5)
    m = TTEmbeddingBag(2449029, 100, [16, 16], [125, 140, 140], [4, 5, 5], optimizer=OptimType.SGD,
                       learning_rate=0.01, sparse=False, use_cache=False, weight_dist="normal").to(DEV)
    g = torch.Generator().manual_seed(7)
    idx = torch.randint(0, 2449029, (40000,), generator=g).to(DEV)
    off = torch.arange(40001, device=DEV)
    want = m(idx, off).detach().clone()
    assert m.prepare(idx, off, 1)
    got = m(idx, off)
    assert torch.equal(got.detach(), want)
    got.sum().backward()
    g1 = [c.grad.clone() for c in m.tt_cores]
    for c in m.tt_cores:
        c.grad = None
    m(idx, off).sum().backward()
    for a, c in zip(g1, m.tt_cores):
        assert float((a - c.grad).abs().max() / c.grad.abs().max()) < 1e-6


# --------------------------------------------------------------------------------------------
# four cores on the three-core kernels (the first two cores merged): run_script.sh:501-541
# --------------------------------------------------------------------------------------------
def test_four_core_recipe_runs_on_the_three_core_kernels(ttg_lib):
    """p = 50,60,60,60 q = 4,2,4,4 ranks 16,16,16 (the reference's final GCN / GAT recipe): forward, dense
    gradients and both fused updates against the oracle, and against this library's shape-generic kernels."""
    import _ttg
    import tt_embeddings as te
    p, q, r, n_emb = [50, 60, 60, 60], [4, 2, 4, 4], [1, 16, 16, 16, 1], 169343
    D = 128
    g = torch.Generator().manual_seed(51)
    cores_cpu = [torch.randn(1, p[t], r[t] * q[t] * r[t + 1], generator=g) * (0.6 / np.sqrt(r[t])) for t in range(4)]
    cn = [c.numpy() for c in cores_cpu]
    rng = np.random.default_rng(7)
    nnz = 30000
    idx = rng.integers(0, n_emb, size=nnz).astype(np.int64)
    idx[:5000] = np.arange(5000)                      # a contiguous stretch, as in a full-graph lookup
    idx[5] = idx[6] = idx[7]
    idx[9] = -3                                       # invalid ids contribute nothing
    row = rng.permutation(nnz).astype(np.int64)
    valid = idx >= 0
    tb = torch.zeros(nnz, dtype=torch.int64, device=DEV)
    dO = ((rng.random(size=(1, nnz, D)) - 0.5) * 0.2).astype(np.float32)
    want = orc.tt_forward(p, q, r, cn, idx[valid], row[valid], nnz)
    wd = orc.tt_backward_dense(p, q, r, cn, idx[valid], row[valid], dO)
    cores = [c.to(DEV) for c in cores_cpu]
    out = te.tt_forward(1000, 1, nnz, D, p, q, r, None, nnz, _t(idx), _t(row), tb, cores)
    assert rel_err(out.cpu().numpy(), want) < TOL
    gd = te.tt_dense_backward(1000, D, p, q, r, None, nnz, _t(idx), _t(row), tb, _t(dO), cores)
    for t in range(4):
        assert rel_err(gd[t].cpu().numpy(), wd[t]) < TOL, "core %d" % t
    te.EXTRA_FLAGS = _ttg.FLAG_FORCE_GENERIC
    try:
        out_g = te.tt_forward(1000, 1, nnz, D, p, q, r, None, nnz, _t(idx), _t(row), tb, cores)
    finally:
        te.EXTRA_FLAGS = 0
    assert float((out - out_g).abs().max() / out_g.abs().max()) < TOL
    cols = [r[t] * q[t] * r[t + 1] for t in range(4)]
    # fused SGD
    dev = [c.to(DEV) for c in cores_cpu]
    te.tt_forward(1000, 1, nnz, D, p, q, r, None, nnz, _t(idx), _t(row), tb, dev)     # leaves its plan behind
    te.tt_sgd_backward(1000, D, 0.1, p, q, r, None, nnz, _t(idx), _t(row), tb, _t(dO), dev)
    ws = [c.copy() for c in cn]
    orc.apply_optimizer(p, cols, "sgd", 0.1, 0.0, ws, None, wd)
    for t in range(4):
        assert rel_err(dev[t].cpu().numpy(), ws[t]) < TOL
    # fused Adagrad, two steps
    dev = [c.to(DEV) for c in cores_cpu]
    state = [torch.zeros_like(c) for c in dev]
    wa = [c.copy() for c in cn]
    wstate = [np.zeros_like(c) for c in wa]
    for _ in range(2):
        te.tt_adagrad_backward(1000, D, 0.05, 1e-10, p, q, r, None, nnz, _t(idx), _t(row), tb, _t(dO), state, dev)
        gr = orc.tt_backward_dense(p, q, r, wa, idx[valid], row[valid], dO)
        orc.apply_optimizer(p, cols, "adagrad", 0.05, 1e-10, wa, wstate, gr)
    for t in range(4):
        assert rel_err(state[t].cpu().numpy(), wstate[t]) < TOL
        assert rel_err(dev[t].cpu().numpy(), wa[t]) < TOL
