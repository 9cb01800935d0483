"""Worker of tests/test_multi_gpu.py (run under torchrun, one rank per GPU): pair sharding through
the C-ABI communicator (cs_b200_comm_*, cs_b200_create_bcast) against a single-GPU run."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import circuitscape_b200 as cb  # noqa: E402
from circuitscape_b200 import dist as cdist, graph  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)

    def exchange(raw):
        t = torch.zeros(128, dtype=torch.uint8, device=dev)
        if raw is not None:
            t = torch.tensor(list(raw), dtype=torch.uint8, device=dev)
        dist.broadcast(t, src=0)
        return bytes(t.cpu().tolist())

    comm = cdist.Comm(local, rank, world, exchange)
    L, _ = graph.synthetic_raster_laplacian(420, 380, seed=13)          # every rank can build it: the root's copy is used
    n, nnz = L.shape[0], L.nnz
    nodes = graph.focal_nodes(n, 7, seed=7)
    src, dst = graph.all_pairs(nodes)
    npairs = len(src)
    solver = cb.CUDASolver(device=local)
    f = comm.create_factor(L if rank == 0 else None, solver, shape=(n, nnz))
    mine = cdist.shard_pairs(npairs, rank, world)
    f.reset_currents()
    out = f.solve_pairs(src[mine], dst[mine], accumulate=True)
    R = comm.gather_pairs(mine, out["R"], npairs)
    comm.reduce_currents(f)
    cum, mx = f.read_currents()
    tmax = comm.max([float(rank), 1.0])
    assert tmax[0] == world - 1 and tmax[1] == 1.0
    # reference: the whole job on this rank's GPU through the plain single-GPU path
    with cb.B200Factor(L, solver) as g:
        ref = g.solve_pairs(src, dst, accumulate=True)
        cref, mref = g.read_currents()
    assert np.abs(R - ref["R"]).max() <= 1e-9 * np.abs(ref["R"]).max(), (rank, R, ref["R"])
    assert np.abs(cum - cref).max() <= 1e-9 * np.abs(cref).max()
    assert np.abs(mx - mref).max() <= 1e-9 * np.abs(mref).max()
    lv_b, lv_s = f.levels(), None
    with cb.B200Factor(L, solver) as g:
        lv_s = g.levels()
    assert len(lv_b) == len(lv_s) and all(a["A"].nnz == b["A"].nnz for a, b in zip(lv_b, lv_s))
    f.close()
    comm.barrier()
    comm.close()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_GPU_OK", world, "ranks", npairs, "pairs")


if __name__ == "__main__":
    main()
