"""Host-side logic of the data-parallel path on CPU: two `gloo` ranks shard the seed nodes of an
epoch disjointly and the single exchange step (one all-reduce over the flat core-gradient buffer)
leaves every rank with the mean gradient.  The optimizer step itself needs CUDA (no CPU path)."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "falcon-ttdforgnns_b200")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import dp
    # products core shapes, rank-dependent gradients
    shapes = [(1, 125, 64), (1, 140, 1280), (1, 140, 80)]
    g = torch.Generator().manual_seed(100 + rank)
    grads = [torch.randn(s, generator=g) for s in shapes]
    reduced = dp.allreduce_mean(grads)
    n = 196615
    perm = dp.epoch_permutation(n, epoch=3, seed=7)
    lo, hi = dp.shard_range(n, rank, world)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), mine=perm[lo:hi].numpy(),
             perm=perm.numpy(), **{"g%d" % i: t.numpy() for i, t in enumerate(grads)},
             **{"r%d" % i: t.numpy() for i, t in enumerate(reduced)})
    dist.destroy_process_group()


def test_two_ranks_exchange_and_shard(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    z = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(world)]
    for i in range(3):
        want = (z[0]["g%d" % i] + z[1]["g%d" % i]) / 2
        for r in range(world):
            np.testing.assert_allclose(z[r]["r%d" % i], want, rtol=0, atol=1e-6)
            assert z[r]["r%d" % i].shape == z[r]["g%d" % i].shape
    # same permutation on both ranks, disjoint shards that cover every seed node
    assert np.array_equal(z[0]["perm"], z[1]["perm"])
    both = np.concatenate([z[0]["mine"], z[1]["mine"]])
    assert both.size == 196615 and np.unique(both).size == 196615
    assert abs(z[0]["mine"].size - z[1]["mine"].size) <= 1


def _merge_worker(rank, world, port, out_dir):
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import dp
    g = torch.Generator().manual_seed(7 + rank)
    keys = torch.randperm(500, generator=g)[:120 + 30 * rank]          # ragged: ranks hold different numbers
    counts = torch.randint(1, 50, (keys.numel(),), generator=g)
    uk, uc = dp.merged_key_counts(keys, counts)
    np.savez(os.path.join(out_dir, "merge%d.npz" % rank), keys=keys.numpy(), counts=counts.numpy(),
             uk=uk.numpy(), uc=uc.numpy())
    dist.destroy_process_group()


def test_two_ranks_merge_lfu_statistics(tmp_path):
    """dp.merged_key_counts: the union of the ranks' keys with the counts of equal keys added, identical on
    every rank (what dp.merge_lfu_statistics writes back into the LFU table before cache_populate)."""
    world = 2
    mp.spawn(_merge_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    z = [np.load(os.path.join(str(tmp_path), "merge%d.npz" % r)) for r in range(world)]
    want = {}
    for r in range(world):
        for k, c in zip(z[r]["keys"], z[r]["counts"]):
            want[int(k)] = want.get(int(k), 0) + int(c)
    ks = np.array(sorted(want))
    for r in range(world):
        assert np.array_equal(z[r]["uk"], ks)
        assert np.array_equal(z[r]["uc"], np.array([want[int(k)] for k in ks]))


def test_merged_key_counts_without_a_process_group():
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    import dp
    uk, uc = dp.merged_key_counts(torch.tensor([5, 3, 5, 9]), torch.tensor([1, 2, 3, 4]))
    assert uk.tolist() == [3, 5, 9] and uc.tolist() == [2, 4, 4]


def test_shard_range_is_balanced_and_complete():
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    import dp
    for n in (0, 1, 7, 193, 196615):
        for world in (1, 2, 4, 8):
            spans = [dp.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
