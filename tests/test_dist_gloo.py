"""N > 1 host logic on CPU: world_size-2 gloo run of the pair sharding, the CSR
broadcast and the resistance gather used by bench.py / dist.py (NCCL on GPUs)."""
import os
import socket

import numpy as np
import pytest
import scipy.sparse as sp

from circuitscape_b200 import dist as cdist
from circuitscape_b200 import graph


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        L = None
        if rank == 0:
            L, _ = graph.synthetic_raster_laplacian(12, 9, seed=3)
        n, nnz, rp, ci, va = cdist.broadcast_csr(L, dist, torch.device("cpu"))
        A = sp.csr_matrix((va.numpy(), ci.numpy(), rp.numpy()), shape=(n, n))
        npairs = 7
        mine = cdist.shard_pairs(npairs, rank, world)
        vals = mine.astype(np.float64) * 10.0 + 1.0          # stand-in for per-pair resistances
        full = cdist.gather_pairs(mine, vals, npairs, dist, device="cpu")
        q.put((rank, n, nnz, float(abs(A).sum()), mine.tolist(), full.tolist()))
    finally:
        dist.destroy_process_group()


def test_two_rank_shard_broadcast_gather():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    L, _ = graph.synthetic_raster_laplacian(12, 9, seed=3)
    for rank, n, nnz, asum, mine, full in res:
        assert (n, nnz) == (L.shape[0], L.nnz)
        assert abs(asum - abs(L).sum()) < 1e-12
        assert full == [i * 10.0 + 1.0 for i in range(7)]
    assert res[0][4] == [0, 2, 4, 6] and res[1][4] == [1, 3, 5]


def test_shard_covers_all_pairs_once():
    for world in (1, 2, 3, 8):
        for npairs in (1, 7, 80, 1000):
            got = np.concatenate([cdist.shard_pairs(npairs, r, world) for r in range(world)])
            assert sorted(got.tolist()) == list(range(npairs))
            sizes = [len(cdist.shard_pairs(npairs, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
