"""Data-parallel host logic on CPU: 2 processes, gloo, the plain-PyTorch backend (N > 1 path of SURVEY.md 8e).
Each rank trains on its own slice of the batch; after one step every replica must hold the same parameters, equal
to a Nesterov step with the AVERAGE of the per-rank gradients (multi_gpu.py:65-69 + training.py:849: the loss is a
mean over the merged batch), with BN batch statistics kept per replica (multi_gpu.py:35-53)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import h_denseunet_b200 as hdn
    from h_denseunet_b200.parallel import DataParallel, shard_bounds
    from torch_backend import TorchBackend
    from util import Args, perturb_params

    def build():
        m = hdn.DenseUNet(reduction=0.5, args=Args(b=1, input_size=64), backend=TorchBackend(), device="cpu", precision="fp32",
                          seed=rank)                      # different initial weights per rank: rank 0's must win
        m.dropout = False
        perturb_params(m, seed=7 + rank)
        m.compile(optimizer=hdn.SGD(lr=1e-3, momentum=0.9, nesterov=True), loss=[hdn.weighted_crossentropy_2ddense])
        return m

    rng = np.random.default_rng(100 + rank)
    x = rng.normal(0, 60, (1, 64, 64, 3)).astype(np.float32)
    y = rng.integers(0, 3, (1, 64, 64, 1)).astype(np.int16)

    m = build()
    hdn.make_parallel(m, world, mini_batch=1)
    assert isinstance(m.dp, DataParallel)
    net = m._net(True)                                    # realises the arenas and broadcasts rank 0's parameters
    w0 = m.params.train.clone()
    ws = [torch.empty_like(w0) for _ in range(world)]
    dist.all_gather(ws, w0)
    assert all(torch.equal(ws[0], w) for w in ws), "replicas do not start from rank 0's parameters"
    # local gradient of this rank's sample (no update), gathered for the expected result
    m.h2d_bytes = m._upload(net, x) + m._labels(net, y)
    net.forward()
    net.backward()
    g_local = m.params.grads.clone()
    gs = [torch.empty_like(g_local) for _ in range(world)]
    dist.all_gather(gs, g_local)
    g_avg = sum(gs) / world
    v = -1e-3 * g_avg
    expect = w0 + 0.9 * v - 1e-3 * g_avg
    # the data-parallel step itself
    m.dp.step(net, 1e-3, 0.9)
    w1 = m.params.train.clone()
    ws = [torch.empty_like(w1) for _ in range(world)]
    dist.all_gather(ws, w1)
    ok_same = all(torch.equal(ws[0], w) for w in ws)
    err = float((w1 - expect).abs().max())
    lo, hi = shard_bounds(m.params.n_train, world, rank)
    out[rank] = (ok_same, err, lo, hi, int(m.params.n_train))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_data_parallel_two_ranks_gloo():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert len(out) == world
    n = out[0][4]
    assert out[0][2] == 0 and out[world - 1][3] == n and out[0][3] == out[1][2], "shards must tile the arena"
    for r in range(world):
        ok_same, err, lo, hi, _ = out[r]
        assert ok_same, "replicas diverged after the data-parallel step"
        assert err < 1e-6, err
        assert lo % 4 == 0


def test_shard_bounds_cover_arena():
    from h_denseunet_b200.parallel import shard_bounds
    for n in (1, 7, 1024, 61_400_003):
        for world in (1, 2, 4, 8):
            prev = 0
            for r in range(world):
                lo, hi = shard_bounds(n, world, r)
                assert lo == prev and lo <= hi <= n and (lo == hi or lo % 4 == 0)
                prev = hi
            assert prev == n
