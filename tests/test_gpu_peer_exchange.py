"""The NVLink peer-memory exchange step (csrc/peer.cu, dp.PeerExchange) against the NCCL path it
replaces: same cores after the same steps, bit-identical replicas.  The two-rank test needs two
GPUs (run with `gpurun --gpus 2`); the one-rank test checks the kernel's arithmetic on one."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "peer_exchange_worker.py")


def _free_port():
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_one_rank_exchange_is_the_plain_update(ttg_lib):
    import torch.distributed as dist
    import dp
    dev = torch.device("cuda", 0)
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % _free_port(), rank=0,
                            world_size=1)
    try:
        g = torch.Generator().manual_seed(3)
        cores = [torch.randn(n, generator=g).to(dev) for n in (8000, 179200, 11200)]
        state = [torch.rand(c.shape, generator=g).to(dev) for c in cores]
        before = [c.clone() for c in cores]
        st0 = [s.clone() for s in state]
        x = dp.PeerExchange(cores)
        for step, optim in enumerate(("sgd", "adagrad", "dense", "sgd")):
            grads = [torch.randn(c.shape, generator=g).to(dev) for c in cores]
            ref_c = [c.clone() for c in cores]
            ref_s = [s.clone() for s in state]
            out = x.step(grads, cores, optim, 0.1, 1e-3, state)
            if optim == "sgd":
                for c, r, gr in zip(cores, ref_c, grads):
                    torch.testing.assert_close(c, r - 0.1 * gr, rtol=1e-6, atol=1e-7)
            elif optim == "adagrad":
                for c, r, gr, s, rs in zip(cores, ref_c, grads, state, ref_s):
                    torch.testing.assert_close(s, rs + gr * gr, rtol=1e-6, atol=1e-7)
                    torch.testing.assert_close(c, r - 0.1 * gr / (torch.sqrt(rs + gr * gr) + 1e-3),
                                               rtol=1e-5, atol=1e-6)
            else:
                assert torch.equal(out, torch.cat(grads))
                for c, r in zip(cores, ref_c):
                    assert torch.equal(c, r)
        assert x.failed_epoch() == 0
        with pytest.raises(RuntimeError):
            dp.PeerExchange([torch.zeros(6, device=dev)])
        x.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("scatter", ["0", "1"], ids=["all_read_all", "reduce_scatter"])
def test_two_ranks_match_the_nccl_path_and_each_other(ttg_lib, scatter):
    """both modes of the exchange kernel (csrc/peer.cu: every rank reads all copies / reduce-scatter + broadcast,
    the default from four ranks on) on two ranks"""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), WORKER]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT,
                       env=dict(os.environ, TTG_PEER_SCATTER=scatter))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "PEER_EXCHANGE_OK" in r.stdout
