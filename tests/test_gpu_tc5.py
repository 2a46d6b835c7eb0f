"""GPU parity tests of the tcgen05 / tensor-memory kernels (csrc/tt_tc5.cu, TTG_FLAG_TCGEN05) against the fp64 oracle
(oracle/tt_oracle.c restates FBTT/tt_embeddings_cuda.cu:967-1081, :421-654) at 1e-5, on the shapes
the kernels are instantiated for, including BASELINE.json's full batch and index range, and against
the other two implementations of the library (FFMA kernels, mma.sync kernels).

Tolerance: 1e-5 of the tensor's largest magnitude against the fp64 oracle (north star).
"""
import numpy as np
import pytest
import torch

from helpers import rel_err
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

TOL = 1e-5
DEV = "cuda:0"
TC5 = 64          # TTG_FLAG_TCGEN05

SHAPES = {
    "products": ([125, 140, 140], [4, 5, 5], [16, 16], 2449029),
    "arxiv": ([55, 55, 56], [4, 4, 8], [16, 16], 169343),
    "cora": ([14, 14, 14], [4, 4, 8], [16, 16], 2708),
}


def _cores(p, q, r, n_emb, seed, num_tables=1):
    g = torch.Generator().manual_seed(seed)
    rr = [1] + list(r) + [1]
    return [torch.randn(num_tables, p[t], rr[t] * q[t] * rr[t + 1], generator=g) / (n_emb ** 0.25)
            for t in range(3)]


@pytest.fixture(params=[64, 1024], ids=["tcgen05", "right_mma"])
def te(ttg_lib, request):
    """the two engines of the right-grouped path: TTG_FLAG_TCGEN05 and TTG_FLAG_RIGHT (mma.sync)"""
    import tt_embeddings
    global TC5
    TC5 = request.param
    tt_embeddings.EXTRA_FLAGS = 0
    yield tt_embeddings
    tt_embeddings.EXTRA_FLAGS = 0


def _fwd(te, shape, cores, idx, row, B, tb=None, num_tables=1, flags=None):
    p, q, r, _ = shape
    te.EXTRA_FLAGS = TC5 if flags is None else flags
    try:
        idx_t = torch.from_numpy(idx).to(DEV)
        row_t = torch.from_numpy(row).to(DEV)
        tb_t = torch.zeros_like(idx_t) if tb is None else torch.from_numpy(tb).to(DEV)
        return te.tt_forward(1000, num_tables, B, int(np.prod(q)), p, q, r, None, idx.size, idx_t, row_t,
                             tb_t, [c.to(DEV) for c in cores])
    finally:
        te.EXTRA_FLAGS = 0


@pytest.mark.parametrize("name,nnz", [("cora", 2708), ("cora", 300), ("arxiv", 20000), ("products", 60000)])
def test_forward_against_oracle(te, name, nnz):
    shape = SHAPES[name]
    p, q, r, n_emb = shape
    cores = _cores(p, q, r, n_emb, 3)
    rng = np.random.default_rng(1)
    # dense in groups: ids from a window that holds about nnz / 3 groups
    hi = min(n_emb, max(nnz, p[1] * p[2] * max(1, nnz // (3 * p[1] * p[2]) + 1)))
    idx = rng.integers(0, hi, size=nnz).astype(np.int64)
    row = np.arange(nnz, dtype=np.int64)
    out = _fwd(te, shape, cores, idx, row, nnz)
    want = orc.tt_forward(p, q, r, [c.numpy() for c in cores], idx, row, nnz)
    assert rel_err(out.cpu().numpy(), want) < TOL


def test_forward_full_size_products_against_oracle_and_other_kernels(te):
    """BASELINE config 2: 262,144 distinct ids over the whole index range."""
    shape = SHAPES["products"]
    p, q, r, n_emb = shape
    cores = _cores(p, q, r, n_emb, 5)
    g = torch.Generator().manual_seed(0)
    nnz = 262144
    idx = torch.randperm(n_emb, generator=g)[:nnz].numpy().astype(np.int64)
    row = np.arange(nnz, dtype=np.int64)
    out = _fwd(te, shape, cores, idx, row, nnz)
    orc.use_all_host_threads()
    want = orc.tt_forward(p, q, r, [c.numpy() for c in cores], idx, row, nnz)
    assert rel_err(out.cpu().numpy(), want) < TOL
    out_ffma = _fwd(te, shape, cores, idx, row, nnz, flags=16)
    out_sync = _fwd(te, shape, cores, idx, row, nnz, flags=0)
    assert float((out - out_ffma).abs().max() / out_ffma.abs().max()) < TOL
    assert float((out - out_sync).abs().max() / out_sync.abs().max()) < TOL
    # permutation equivariance and sorted input
    perm = torch.randperm(nnz, generator=g).numpy()
    out_p = _fwd(te, shape, cores, np.ascontiguousarray(idx[perm]), row, nnz)
    assert torch.equal(out_p[0], out[0][torch.from_numpy(perm).to(DEV)])


def test_forward_bags_tables_and_invalid_indices(te):
    """several indices per bag (accumulation), empty bags (zero rows), two tables, out-of-range ids
    (skipped, as the sorted kernels do), rows in arbitrary order"""
    shape = SHAPES["cora"]
    p, q, r, n_emb = shape
    cores = _cores(p, q, r, n_emb, 9, num_tables=2)
    rng = np.random.default_rng(4)
    B, nnz = 700, 4000
    idx = rng.integers(0, n_emb, size=nnz).astype(np.int64)
    row = rng.integers(0, B - 50, size=nnz).astype(np.int64)      # the last 50 bags stay empty
    tb = rng.integers(0, 2, size=nnz).astype(np.int64)
    out = _fwd(te, shape, cores, idx, row, B, tb=tb, num_tables=2)
    want = orc.tt_forward(p, q, r, [c.numpy() for c in cores], idx, row, B, tableidx=tb, num_tables=2)
    assert rel_err(out.cpu().numpy(), want) < TOL
    assert float(out[:, B - 50:].abs().max()) == 0.0
    # invalid ids contribute nothing
    idx_bad = idx.copy()
    idx_bad[::7] = 14 ** 3 + 5
    keep = np.ones(nnz, dtype=bool)
    keep[::7] = False
    out_bad = _fwd(te, shape, cores, idx_bad, row, B, tb=tb, num_tables=2)
    want_bad = orc.tt_forward(p, q, r, [c.numpy() for c in cores], idx[keep], row[keep], B,
                              tableidx=tb[keep], num_tables=2)
    assert rel_err(out_bad.cpu().numpy(), want_bad) < TOL


def test_forward_one_giant_group_and_many_tiles(te):
    """one group with 40,000 rows (more than one round of the kernel's tile list: 512 tiles of 32
    rows) next to ordinary groups, duplicates included"""
    shape = SHAPES["cora"]
    p, q, r, n_emb = shape
    cores = _cores(p, q, r, n_emb, 11)
    rng = np.random.default_rng(6)
    hp = p[1] * p[2]
    giant = (rng.integers(0, p[0], size=40000) * hp + 77).astype(np.int64)   # all in group h = 77
    rest = rng.integers(0, n_emb, size=5000).astype(np.int64)
    idx = np.concatenate([giant, rest])
    rng.shuffle(idx)
    nnz = idx.size
    row = np.arange(nnz, dtype=np.int64)
    out = _fwd(te, shape, cores, idx, row, nnz)
    want = orc.tt_forward(p, q, r, [c.numpy() for c in cores], idx, row, nnz)
    assert rel_err(out.cpu().numpy(), want) < TOL


def test_forward_tf32_mode_stated_bound(te):
    shape = SHAPES["arxiv"]
    p, q, r, n_emb = shape
    cores = _cores(p, q, r, n_emb, 13)
    rng = np.random.default_rng(8)
    nnz = 30000
    idx = rng.integers(0, n_emb, size=nnz).astype(np.int64)
    row = np.arange(nnz, dtype=np.int64)
    out = _fwd(te, shape, cores, idx, row, nnz, flags=TC5 | 8)
    want = orc.tt_forward(p, q, r, [c.numpy() for c in cores], idx, row, nnz)
    assert rel_err(out.cpu().numpy(), want) < 3e-3


def _bwd(te, shape, cores, idx, row, dO, tb=None, flags=None, mode="dense", lr=0.1, state=None):
    p, q, r, _ = shape
    D = int(np.prod(q))
    te.EXTRA_FLAGS = TC5 if flags is None else flags
    try:
        idx_t = torch.from_numpy(idx).to(DEV)
        row_t = torch.from_numpy(row).to(DEV)
        tb_t = torch.zeros_like(idx_t) if tb is None else torch.from_numpy(tb).to(DEV)
        dO_t = torch.from_numpy(dO).to(DEV)
        if mode == "dense":
            return te.tt_dense_backward(1000, D, p, q, r, None, idx.size, idx_t, row_t, tb_t, dO_t, cores)
        if mode == "sgd":
            return te.tt_sgd_backward(1000, D, lr, p, q, r, None, idx.size, idx_t, row_t, tb_t, dO_t, cores)
        return te.tt_adagrad_backward(1000, D, lr, 1e-10, p, q, r, None, idx.size, idx_t, row_t, tb_t, dO_t,
                                      state, cores)
    finally:
        te.EXTRA_FLAGS = 0


@pytest.mark.parametrize("name,nnz", [("cora", 2708), ("cora", 300), ("arxiv", 20000), ("products", 60000)])
def test_dense_backward_against_oracle(te, name, nnz):
    shape = SHAPES[name]
    p, q, r, n_emb = shape
    D = int(np.prod(q))
    cores = _cores(p, q, r, n_emb, 3)
    rng = np.random.default_rng(2)
    hi = min(n_emb, max(nnz, p[1] * p[2] * max(1, nnz // (3 * p[1] * p[2]) + 1)))
    idx = rng.integers(0, hi, size=nnz).astype(np.int64)
    row = rng.permutation(nnz).astype(np.int64)
    dO = ((rng.random(size=(1, nnz, D)) - 0.5) * 0.2).astype(np.float32)
    got = _bwd(te, shape, [c.to(DEV) for c in cores], idx, row, dO)
    want = orc.tt_backward_dense(p, q, r, [c.numpy() for c in cores], idx, row, dO)
    for t in range(3):
        assert rel_err(got[t].cpu().numpy(), want[t]) < TOL, "core %d" % t


def test_backward_full_size_products_against_oracle(te):
    """BASELINE config 2: 262,144 distinct ids over the whole index range; forward first, so the backward
    runs on the forward's plan and table (TTG_FLAG_PLAN_VALID), then once more on its own."""
    shape = SHAPES["products"]
    p, q, r, n_emb = shape
    D = 100
    cores = _cores(p, q, r, n_emb, 5)
    dcores = [c.to(DEV) for c in cores]
    g = torch.Generator().manual_seed(0)
    nnz = 262144
    idx = torch.randperm(n_emb, generator=g)[:nnz].numpy().astype(np.int64)
    row = np.arange(nnz, dtype=np.int64)
    dO = ((torch.rand(1, nnz, D, generator=g) - 0.5) * 0.2).numpy()
    idx_t, row_t = torch.from_numpy(idx).to(DEV), torch.from_numpy(row).to(DEV)
    tb_t, dO_t = torch.zeros_like(idx_t), torch.from_numpy(dO).to(DEV)
    te.EXTRA_FLAGS = TC5
    te.tt_forward(1000, 1, nnz, D, p, q, r, None, nnz, idx_t, row_t, tb_t, dcores)
    got = te.tt_dense_backward(1000, D, p, q, r, None, nnz, idx_t, row_t, tb_t, dO_t, dcores)
    te.EXTRA_FLAGS = 0
    orc.use_all_host_threads()
    want = orc.tt_backward_dense(p, q, r, [c.numpy() for c in cores], idx, row, dO)
    for t in range(3):
        assert rel_err(got[t].cpu().numpy(), want[t]) < TOL, "core %d" % t
    got2 = _bwd(te, shape, dcores, idx, row, dO)
    for t in range(3):
        assert rel_err(got2[t].cpu().numpy(), want[t]) < TOL, "core %d (own plan)" % t
    # the two other implementations of the library agree
    for fl in (16, 0):
        other = _bwd(te, shape, dcores, idx, row, dO, flags=fl)
        for t in range(3):
            assert rel_err(got[t].cpu().numpy(), other[t].cpu().numpy()) < TOL


def test_backward_bags_tables_duplicates_and_giant_group(te):
    shape = SHAPES["cora"]
    p, q, r, n_emb = shape
    D = 128
    cores = _cores(p, q, r, n_emb, 9, num_tables=2)
    rng = np.random.default_rng(4)
    hp = p[1] * p[2]
    giant = (rng.integers(0, p[0], size=30000) * hp + 77).astype(np.int64)   # one group, heavy duplicates
    rest = rng.integers(0, n_emb, size=6000).astype(np.int64)
    idx = np.concatenate([giant, rest])
    rng.shuffle(idx)
    nnz, B = idx.size, 5000
    row = rng.integers(0, B, size=nnz).astype(np.int64)
    tb = rng.integers(0, 2, size=nnz).astype(np.int64)
    dO = ((rng.random(size=(2, B, D)) - 0.5) * 0.2).astype(np.float32)
    got = _bwd(te, shape, [c.to(DEV) for c in cores], idx, row, dO, tb=tb)
    want = orc.tt_backward_dense(p, q, r, [c.numpy() for c in cores], idx, row, dO, tableidx=tb, num_tables=2)
    for t in range(3):
        assert rel_err(got[t].cpu().numpy(), want[t]) < TOL, "core %d" % t


def test_fused_sgd_and_adagrad_updates(te):
    shape = SHAPES["arxiv"]
    p, q, r, n_emb = shape
    D = 128
    rr = [1] + list(r) + [1]
    cols = [rr[t] * q[t] * rr[t + 1] for t in range(3)]
    cores = _cores(p, q, r, n_emb, 17)
    rng = np.random.default_rng(5)
    nnz = 12000
    idx = rng.integers(0, n_emb, size=nnz).astype(np.int64)
    row = np.arange(nnz, dtype=np.int64)
    dO = ((rng.random(size=(1, nnz, D)) - 0.5) * 0.2).astype(np.float32)
    grads = orc.tt_backward_dense(p, q, r, [c.numpy() for c in cores], idx, row, dO)
    dev = [c.to(DEV) for c in cores]
    _bwd(te, shape, dev, idx, row, dO, mode="sgd", lr=0.1)
    want = [c.numpy().copy() for c in cores]
    orc.apply_optimizer(p, cols, "sgd", 0.1, 0.0, want, None, grads)
    for t in range(3):
        assert rel_err(dev[t].cpu().numpy(), want[t]) < TOL
    dev = [c.to(DEV) for c in cores]
    state = [torch.zeros_like(c) for c in dev]
    want = [c.numpy().copy() for c in cores]
    wstate = [np.zeros_like(c) for c in want]
    for _ in range(2):
        _bwd(te, shape, dev, idx, row, dO, mode="adagrad", lr=0.05, state=state)
        g = orc.tt_backward_dense(p, q, r, want, idx, row, dO)
        orc.apply_optimizer(p, cols, "adagrad", 0.05, 1e-10, want, wstate, g)
    for t in range(3):
        assert rel_err(state[t].cpu().numpy(), wstate[t]) < TOL
        assert rel_err(dev[t].cpu().numpy(), want[t]) < TOL


# --------------------------------------------------------------------------------------------
# adversarial index distributions for the right-grouped kernels (run ownership, tails, invalid keys)
# --------------------------------------------------------------------------------------------
def _both_ways(te, shape, idx, B=None, row=None, seed=1):
    p, q, r, n_emb = shape
    D = int(np.prod(q))
    cores = _cores(p, q, r, n_emb, 31)
    cn = [c.numpy() for c in cores]
    nnz = idx.size
    row = np.arange(nnz, dtype=np.int64) if row is None else row
    B = nnz if B is None else B
    valid = (idx >= 0) & (idx < int(np.prod(p)))
    out = _fwd(te, shape, cores, idx, row, B)
    want = orc.tt_forward(p, q, r, cn, idx[valid], row[valid], B)
    assert rel_err(out.cpu().numpy(), want) < TOL
    dO = (np.random.default_rng(seed).random(size=(1, B, D)).astype(np.float32) - 0.5) * 0.2
    got = _bwd(te, shape, [c.to(DEV) for c in cores], idx, row, dO)
    wd = orc.tt_backward_dense(p, q, r, cn, idx[valid], row[valid], dO)
    for t in range(3):
        assert rel_err(got[t].cpu().numpy(), wd[t]) < TOL, "core %d" % t


@pytest.mark.parametrize("name", ["products", "arxiv"])
def test_one_right_group_holds_the_batch(te, name):
    """Every row in one (i1, i2) group -- one warp owns the group's S1, the runs of all other warps are empty --
    then the same group under a uniform background, and a batch that is one id repeated."""
    shape = SHAPES[name]
    p, q, r, n_emb = shape
    rng = np.random.default_rng(17)
    hp = p[1] * p[2]
    h = 37 * p[2] + 5
    i0_max = (n_emb - 1 - h) // hp
    one_group = (h + hp * rng.integers(0, i0_max + 1, size=6000)).astype(np.int64)
    _both_ways(te, shape, one_group)
    mixed = np.concatenate([one_group[:3000], np.full(1000, h + hp * 3),
                            rng.integers(0, n_emb, size=max(hp, 6000))]).astype(np.int64)
    rng.shuffle(mixed)
    _both_ways(te, shape, mixed)
    _both_ways(te, shape, np.full(hp + 7, h + hp * 2, dtype=np.int64))


def test_ragged_batches_around_the_group_count(te):
    shape = SHAPES["cora"]           # 196 (i1, i2) groups: the right-grouped path starts at nnz = 196
    rng = np.random.default_rng(23)
    for nnz in (196, 197, 211, 255, 513, 1000):
        _both_ways(te, shape, rng.integers(0, shape[3], size=nnz).astype(np.int64), seed=nnz)


def test_invalid_indices_and_bags_backward(te):
    """Out-of-range / negative ids sort behind every valid key and contribute nothing to either pass; bags of
    0..7 indices share one d_output row."""
    shape = SHAPES["arxiv"]
    p, q, r, n_emb = shape
    rng = np.random.default_rng(29)
    lengths = rng.integers(0, 8, size=3000)
    nnz = int(lengths.sum())
    idx = rng.integers(0, n_emb, size=nnz).astype(np.int64)
    bad = rng.permutation(nnz)[:200]
    idx[bad[:100]] = -1 - rng.integers(0, 5, size=100)
    idx[bad[100:]] = int(np.prod(p)) + rng.integers(0, 1000, size=100)
    row = np.repeat(np.arange(lengths.size), lengths).astype(np.int64)
    _both_ways(te, shape, idx, B=lengths.size, row=row)


def test_large_calls_take_the_right_grouped_kernels_by_themselves(ttg_lib):
    """From 14 rows per (i1, i2) group and call on (274,400 rows at products shape), a call without engine flags runs the right-grouped mma.sync kernels
    (tt_sorted.cu use_r): bit-identical to the same call with TTG_FLAG_RIGHT, 1e-5 from the left-grouped kernels
    (TTG_FLAG_MMA_SYNC); below that size it is bit-identical to the left-grouped ones."""
    import _ttg
    import tt_embeddings as te
    shape = SHAPES["products"]
    p, q, r, n_emb = shape
    cores = [c.to(DEV) for c in _cores(p, q, r, n_emb, 9)]
    g = torch.Generator().manual_seed(4)
    for nnz, same_as in ((393216, _ttg.FLAG_RIGHT), (131072, _ttg.FLAG_MMA_SYNC)):
        idx = torch.randperm(n_emb, generator=g)[:nnz].to(DEV)
        row = torch.arange(nnz, device=DEV)
        tb = torch.zeros_like(idx)
        dO = ((torch.rand(1, nnz, 100, generator=g) - 0.5) * 0.2).to(DEV)
        res = {}
        for fl in (0, _ttg.FLAG_RIGHT, _ttg.FLAG_MMA_SYNC):
            te.EXTRA_FLAGS = fl
            try:
                out = te.tt_forward(1000, 1, nnz, 100, p, q, r, None, nnz, idx, row, tb, cores)
                gr = te.tt_dense_backward(1000, 100, p, q, r, None, nnz, idx, row, tb, dO, cores)
            finally:
                te.EXTRA_FLAGS = 0
            res[fl] = [out] + [x.clone() for x in gr]
        other = _ttg.FLAG_MMA_SYNC if same_as == _ttg.FLAG_RIGHT else _ttg.FLAG_RIGHT
        assert torch.equal(res[0][0], res[same_as][0])      # forward: no atomics on either path
        assert not torch.equal(res[0][0], res[other][0])
        for a, b in zip(res[0], res[same_as]):              # gradients: one core per path goes through atomics
            assert float((a - b).abs().max() / b.abs().max()) < TOL
        for a, b in zip(res[0], res[other]):
            assert float((a - b).abs().max() / b.abs().max()) < TOL


def test_right_grouped_kernels_on_the_deterministic_plan(te):
    """TTG_FLAG_DETERMINISTIC (radix-sorted plan, fixed order inside a group) together with the right-grouped
    engines: same results at 1e-5, and the forward twice bit-identical."""
    import _ttg
    shape = SHAPES["arxiv"]
    p, q, r, n_emb = shape
    cores = _cores(p, q, r, n_emb, 13)
    rng = np.random.default_rng(6)
    nnz = 40000
    idx = rng.integers(0, n_emb, size=nnz).astype(np.int64)
    row = rng.permutation(nnz).astype(np.int64)
    fl = TC5 | _ttg.FLAG_DETERMINISTIC
    out = _fwd(te, shape, cores, idx, row, nnz, flags=fl)
    out2 = _fwd(te, shape, cores, idx, row, nnz, flags=fl)
    assert torch.equal(out, out2)
    want = orc.tt_forward(p, q, r, [c.numpy() for c in cores], idx, row, nnz)
    assert rel_err(out.cpu().numpy(), want) < TOL
    dO = ((rng.random(size=(1, nnz, int(np.prod(q)))) - 0.5) * 0.2).astype(np.float32)
    got = _bwd(te, shape, [c.to(DEV) for c in cores], idx, row, dO, flags=fl)
    wd = orc.tt_backward_dense(p, q, r, [c.numpy() for c in cores], idx, row, dO)
    for t in range(3):
        assert rel_err(got[t].cpu().numpy(), wd[t]) < TOL, "core %d" % t
