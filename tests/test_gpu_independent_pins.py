"""Rows a-9 / f-1 of SURVEY section 8 are restated from DGL 2.1's documentation (DGL is not in this image, so
those rows are parity-unpinned against the reference's dependency).  This file pins them against INDEPENDENT
implementations instead: scipy.sparse CSR products and torch.sparse.mm in float64 for the aggregation and the
SAGEConv / GraphConv layers (written from the formulas in DGL's docstrings, not from this repo's oracle), and
graph-theoretic invariants checked with numpy for the sampler's blocks.  Tolerance 1e-5 of the tensor maximum.
"""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-5


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def _random_block(rng, num_src, num_dst, max_deg):
    deg = rng.integers(0, max_deg + 1, size=num_dst)
    deg[::17] = 0                      # destination nodes without in-edges
    indptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int64)
    indices = rng.integers(0, num_src, size=int(indptr[-1])).astype(np.int32)   # repeated edges allowed
    return indptr, indices


@pytest.mark.parametrize("F", [100, 128, 47])
@pytest.mark.parametrize("mean", [True, False])
def test_aggregation_against_scipy_and_torch_sparse(ttg_lib, F, mean):
    import gnn_ops
    rng = np.random.default_rng(F + mean)
    num_src, num_dst = 5000, 1700
    indptr, indices = _random_block(rng, num_src, num_dst, 12)
    x = rng.standard_normal((num_src, F)).astype(np.float32)
    A = sp.csr_matrix((np.ones(indices.size), indices.astype(np.int64), indptr), shape=(num_dst, num_src))
    deg = np.maximum(np.diff(indptr), 1).astype(np.float64)
    want = A @ x.astype(np.float64)
    if mean:
        want = want / deg[:, None]
    block = gnn_ops.Block(torch.from_numpy(indptr).to(DEV), torch.from_numpy(indices).to(DEV), num_src, num_dst)
    xt = torch.from_numpy(x).to(DEV).requires_grad_(True)
    out = gnn_ops.aggregate(block, xt, mean=mean)
    assert _rel(out.detach().cpu().numpy(), want) < TOL
    # backward: d x = A^T (d out / deg) -- torch.sparse in float64 as a second opinion
    dout = rng.standard_normal((num_dst, F)).astype(np.float32)
    out.backward(torch.from_numpy(dout).to(DEV))
    At = torch.sparse_coo_tensor(
        torch.from_numpy(np.stack([indices.astype(np.int64), np.repeat(np.arange(num_dst), np.diff(indptr))])),
        torch.ones(indices.size, dtype=torch.float64), (num_src, num_dst)).coalesce()
    d = torch.from_numpy(dout).double()
    if mean:
        d = d / torch.from_numpy(deg)[:, None]
    want_dx = torch.sparse.mm(At, d).numpy()
    assert _rel(xt.grad.cpu().numpy(), want_dx) < TOL


def test_sage_conv_and_graph_conv_against_dense_float64(ttg_lib):
    """DGL docstrings: SAGEConv(mean): h_v = W_self h_v + W_neigh mean_{u in N(v)} h_u + b (the neighbour
    transform first when in > out); GraphConv(norm='both'): h = D_in^-1/2 A D_out^-1/2 X W + b, degrees
    clamped to 1."""
    import gnn_ops
    rng = np.random.default_rng(11)
    num_src, num_dst = 900, 400
    indptr, indices = _random_block(rng, num_src, num_dst, 9)
    A = sp.csr_matrix((np.ones(indices.size), indices.astype(np.int64), indptr), shape=(num_dst, num_src)).toarray()
    block = gnn_ops.Block(torch.from_numpy(indptr).to(DEV), torch.from_numpy(indices).to(DEV), num_src, num_dst)
    for fin, fout in ((100, 256), (256, 47)):
        torch.manual_seed(fin)
        layer = gnn_ops.SAGEConv(fin, fout, "mean").to(DEV)
        x = torch.randn(num_src, fin, device=DEV)
        out = layer(block, (x, x[:num_dst]))
        Ws, Wn, b = (layer.fc_self.weight.detach().cpu().double().numpy(),
                     layer.fc_neigh.weight.detach().cpu().double().numpy(),
                     layer.bias.detach().cpu().double().numpy())
        xd = x.cpu().double().numpy()
        deg = np.maximum(A.sum(1), 1.0)
        want = xd[:num_dst] @ Ws.T + ((A @ xd) / deg[:, None]) @ Wn.T + b
        assert _rel(out.detach().cpu().numpy(), want) < 3e-5     # fp32 GEMMs of the dense part (torch library)
        gc = gnn_ops.GraphConv(fin, fout).to(DEV)
        out = gc(block, x)
        W, b = gc.weight.detach().cpu().double().numpy(), gc.bias.detach().cpu().double().numpy()
        dout_deg = np.maximum(A.sum(0), 1.0)
        din_deg = np.maximum(A.sum(1), 1.0)
        want = ((A @ (xd / np.sqrt(dout_deg)[:, None])) / np.sqrt(din_deg)[:, None]) @ W + b
        assert _rel(out.detach().cpu().numpy(), want) < 3e-5


def test_sampled_blocks_satisfy_the_neighbor_sampler_contract(ttg_lib):
    """dgl.dataloading.NeighborSampler contract (DGL user guide 6.1): every sampled edge is an edge of the graph,
    a node with degree <= fanout keeps ALL its in-edges, otherwise exactly `fanout` distinct ones; destination
    nodes are the first source nodes; input_nodes are unique global ids; the seeds come back as output_nodes."""
    import sage
    import sampler
    dev = torch.device(DEV)
    g = sage.synthetic_graph(20000, 300000, dev, seed=5)
    indptr, nbrs = g.indptr.cpu().numpy(), g.indices.cpu().numpy()
    smp = sampler.NeighborSampler([3, 5])
    seeds = torch.randperm(20000, generator=torch.Generator().manual_seed(1))[:512].to(dev)
    inp, outp, blocks = smp.sample_blocks(g, seeds, seed=9)
    assert torch.equal(outp, seeds)
    inp_np = inp.cpu().numpy()
    assert np.unique(inp_np).size == inp_np.size
    src_ids = inp_np
    for blk, fanout in zip(blocks, [3, 5]):
        bp, bi = blk.indptr.cpu().numpy(), blk.indices.cpu().numpy()
        assert blk.num_src == src_ids.size
        dst_ids = src_ids[:blk.num_dst]                      # destination nodes first
        for v in range(0, blk.num_dst, 7):
            gv = dst_ids[v]
            full = nbrs[indptr[gv]:indptr[gv + 1]]
            got = src_ids[bi[bp[v]:bp[v + 1]]]
            deg = full.size
            if deg <= fanout:
                assert sorted(got.tolist()) == sorted(full.tolist())
            else:
                assert got.size == fanout
                # distinct positions of the adjacency list: as a multiset, a sub-multiset of it
                fl = full.tolist()
                for u in got.tolist():
                    assert u in fl
                    fl.remove(u)
        src_ids = dst_ids
    assert np.array_equal(src_ids, seeds.cpu().numpy())
