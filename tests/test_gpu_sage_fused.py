"""SAGE with fuse_input: the first layer's neighbour mean taken by the TT lookup itself (bags = destination
nodes followed by their sampled neighbours) against the plain path that reconstructs all num_src rows and
aggregates them (gnn_model.py:199-217).  Same weights, dropout off: logits, the dense layers' gradients and
the cores after the fused SGD step must agree to 1e-5 (relative to the tensor's maximum)."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-5


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


def _models(n_nodes, sparse):
    import sage
    torch.manual_seed(3)
    plain = sage.SAGE(n_nodes, 100, 128, 7, 3, 0.0, (16, 16), (28, 28, 28), (4, 5, 5), sparse=sparse,
                      learning_rate=0.05, embed_name="fbtt", fuse_input=False).to(DEV)
    fused = sage.SAGE(n_nodes, 100, 128, 7, 3, 0.0, (16, 16), (28, 28, 28), (4, 5, 5), sparse=sparse,
                      learning_rate=0.05, embed_name="fbtt", fuse_input=True).to(DEV)
    fused.load_state_dict(copy.deepcopy(plain.state_dict()))
    with torch.no_grad():            # gradients well above the noise floor of the comparison
        for m in (plain, fused):
            for c in m.embed_layer.tt_cores:
                c.mul_(8.0)
    assert fused.fuse_input and not plain.fuse_input
    return plain, fused


@pytest.mark.parametrize("sparse", [True, False])
def test_fused_first_layer_equals_plain_path(ttg_lib, sparse):
    import sage
    import sampler
    n_nodes = 20000
    g = sage.synthetic_graph(n_nodes, 300000, torch.device(DEV), seed=5)
    smp = sampler.NeighborSampler([4, 6, 8])
    seeds = torch.randperm(n_nodes, generator=torch.Generator().manual_seed(2))[:512].to(DEV)
    inp, outp, blocks = smp.sample_blocks(g, seeds, seed=11)
    labels = torch.randint(0, 7, (outp.numel(),), generator=torch.Generator().manual_seed(4)).to(DEV)
    plain, fused = _models(n_nodes, sparse)
    before = [c.detach().clone() for c in plain.embed_layer.tt_cores]
    outs = []
    for m in (plain, fused):
        m.train()
        logits = m(blocks, inp)
        loss = torch.nn.functional.cross_entropy(logits, labels)
        loss.backward()
        outs.append(logits.detach())
    torch.cuda.synchronize()
    assert rel(outs[1], outs[0]) < TOL
    for a, b in zip(fused.dense_parameters(), plain.dense_parameters()):
        assert rel(a.grad, b.grad) < TOL
    for t, (a, b) in enumerate(zip(fused.embed_layer.tt_cores, plain.embed_layer.tt_cores)):
        if sparse:
            # the cores were updated inside the backward (core - lr * g, rounded to fp32: an ulp of the core is
            # about 1e-3 of the largest update here, so the updates themselves can only be compared that coarsely;
            # the gradients are compared at 1e-5 in the sparse=False case)
            assert rel(a.detach(), b.detach()) < 1e-6
            assert rel(a.detach() - before[t], b.detach() - before[t]) < 5e-3
        else:
            assert rel(a.grad, b.grad) < TOL


def test_fused_first_layer_handles_destinations_without_neighbours(ttg_lib):
    """A destination whose sample is empty contributes a zero mean (DGL's mean over no messages)."""
    import sage
    from gnn_ops import Block
    plain, fused = _models(5000, True)
    indptr = torch.tensor([0, 2, 2, 5], dtype=torch.int64, device=DEV)
    indices = torch.tensor([3, 4, 0, 5, 6], dtype=torch.int32, device=DEV)
    blk = Block(indptr, indices, 7, 3)
    inp = torch.tensor([10, 4000, 77, 1234, 4999, 0, 31], dtype=torch.int64, device=DEV)
    with torch.no_grad():
        for m in (plain, fused):
            m.eval()
        h0 = plain.embed_layer(inp, torch.arange(8, device=DEV))
        want = plain.layers[0](blk, (h0, h0[:3]))
        got = fused._fused_first_layer(blk, inp)
    assert rel(got, want) < TOL
