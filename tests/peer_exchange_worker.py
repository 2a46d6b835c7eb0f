"""Worker of tests/test_gpu_peer_exchange.py (one process per GPU under torchrun): two replicas
of the same TT table per rank, one trained through dp.PeerExchange, one through the NCCL
all-reduce + ttg_apply_optimizer path, on different batches per rank."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "falcon-ttdforgnns_b200")]
import dp  # noqa: E402
from FBTT.tt_embeddings_ops import OptimType, TTEmbeddingBag  # noqa: E402


def main():
    rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n_emb, D, ranks, p, q = 2708, 128, [16, 16], [14, 14, 14], [4, 4, 8]
    for optimizer, name in ((OptimType.SGD, "sgd"), (OptimType.EXACT_ADAGRAD, "adagrad")):
        mods = []
        for _ in range(2):
            torch.manual_seed(0)
            np.random.seed(0)
            m = TTEmbeddingBag(n_emb, D, ranks, p, q, optimizer=optimizer, learning_rate=0.05,
                               sparse=False, use_cache=False, weight_dist="normal")
            with torch.no_grad():
                for c in m.tt_cores:
                    c.mul_(20.0)
            mods.append(m)
        m_peer, m_nccl = mods
        init = {id(c): c.detach().clone() for c in m_peer.tt_cores}
        xchg = dp.PeerExchange(m_peer.tt_cores)
        rng = np.random.default_rng(100 + rank)
        nb = 500
        offsets = torch.arange(nb + 1, device=dev)
        for step in range(7):
            idx = torch.from_numpy(rng.integers(0, n_emb, size=nb)).to(dev)
            tgt = torch.from_numpy(rng.standard_normal((nb, D)).astype(np.float32)).to(dev) * 0.01
            for m, ex in ((m_peer, xchg), (m_nccl, None)):
                (m(idx, offsets) * tgt).sum().backward()
                dp.dp_backward_step(m, [c.grad for c in m.tt_cores], exchange=ex)
                for c in m.tt_cores:
                    c.grad = None
        assert xchg.failed_epoch() == 0
        for a, b in zip(m_peer.tt_cores, m_nccl.tt_cores):
            torch.testing.assert_close(a, b, rtol=2e-5, atol=2e-5 * float(b.abs().max()))
            assert not torch.equal(a, init[id(a)]), "the cores were never updated"
        # replicas are bit-identical: every rank added the copies in the same order
        for c in m_peer.tt_cores:
            gathered = [torch.empty_like(c) for _ in range(world)]
            dist.all_gather(gathered, c.data.contiguous())
            for gth in gathered:
                assert torch.equal(gth, c.data), "replicas diverged (%s)" % name
        # dense mode: the mean gradient itself
        grads = [torch.full_like(c, float(rank + 1)) for c in m_peer.tt_cores]
        mean = xchg.step(grads, m_peer.tt_cores, "dense")
        assert torch.equal(mean, torch.full_like(mean, (world + 1) / 2.0))
        xchg.close()
    # the whole step (forward, backward, exchange + update) replayed as a CUDA graph
    import pipeline
    mods = []
    for _ in range(2):
        torch.manual_seed(0)
        np.random.seed(0)
        mods.append(TTEmbeddingBag(n_emb, D, ranks, p, q, optimizer=OptimType.SGD, learning_rate=0.05,
                                   sparse=False, use_cache=False, weight_dist="normal"))
    m_graph, m_eager = mods
    xchg = dp.PeerExchange(m_graph.tt_cores)
    rng = np.random.default_rng(200 + rank)
    nb = 400
    offsets = torch.arange(nb + 1, device=dev)
    tgt = torch.from_numpy(rng.standard_normal((nb, D)).astype(np.float32)).to(dev) * 0.01
    batches = [torch.from_numpy(rng.integers(0, n_emb, size=nb)).to(dev) for _ in range(5)]
    idx_static = batches[0].clone()

    def step(m, idx, ex):
        (m(idx, offsets) * tgt).sum().backward()
        dp.dp_backward_step(m, [c.grad for c in m.tt_cores], exchange=ex)
        for c in m.tt_cores:
            c.grad = None

    gs = pipeline.GraphedStep(lambda: step(m_graph, idx_static, xchg), dev, warmup=2)
    for _ in range(2):
        step(m_eager, batches[0], None)
    for b in batches[1:]:
        idx_static.copy_(b)
        gs()
        step(m_eager, b, None)
    assert xchg.failed_epoch() == 0
    for a, b in zip(m_graph.tt_cores, m_eager.tt_cores):
        torch.testing.assert_close(a, b, rtol=2e-5, atol=2e-5 * float(b.abs().max()))
    del gs
    xchg.close()
    # the GraphSAGE trainer at N > 1: flat dense-gradient buffer + peer exchange for the cores;
    # replicas stay bit-identical and equal the NCCL-only trainer
    import sage
    import sampler
    graph = sage.synthetic_graph(5000, 60000, dev, seed=0)
    labels = torch.randint(0, 7, (5000,), device=dev, generator=torch.Generator(device=dev).manual_seed(1))
    trainers = []
    for peer in (True, False):
        torch.manual_seed(0)
        np.random.seed(0)
        model = sage.SAGE(5000, 128, 32, 7, 2, 0.0, (16, 16), (14, 18, 20), (4, 4, 8), sparse=False,
                          learning_rate=0.01).to(dev)
        for prm in model.parameters():
            dist.broadcast(prm.data, 0)
        trainers.append(sage.Trainer(model, lr=0.01, world=world, peer_exchange=peer))
    assert trainers[0].xchg is not None and trainers[1].xchg is None
    smp = sampler.NeighborSampler([4, 4])
    for step_i in range(4):
        seeds = torch.randperm(5000, generator=torch.Generator().manual_seed(10 * step_i + rank))[:64].to(dev)
        inp, outp, blocks = smp.sample_blocks(graph, seeds, seed=step_i * world + rank)
        for tr in trainers:
            tr.step(blocks, inp, labels[outp])
    for pa, pb in zip(trainers[0].model.parameters(), trainers[1].model.parameters()):
        torch.testing.assert_close(pa, pb, rtol=1e-4, atol=1e-5 * float(pb.abs().max()) + 1e-7)
        gathered = [torch.empty_like(pa.data) for _ in range(world)]
        dist.all_gather(gathered, pa.data.contiguous())
        for gth in gathered:
            assert torch.equal(gth, pa.data), "SAGE replicas diverged"
    for tr in trainers:
        tr.close()
    dist.barrier()
    if rank == 0:
        print("PEER_EXCHANGE_OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
