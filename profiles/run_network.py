"""Config C5: network-mode power-law graph, advanced "all-to-one".

    python profiles/run_network.py --nodes 2000000 --focal 64            # one GPU
    torchrun --nproc-per-node N ... profiles/run_network.py ...          # columns sharded

Every all-to-one iteration (ground one focal node, 1 A into every other one) is a column
of ONE batched solve on the singular Laplacian (core.all_to_one_batched); ranks take
columns rank::world of the same broadcast operator.  Prints one JSON line on rank 0."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nodes", type=float, default=2e6)
    ap.add_argument("--m", type=int, default=5)
    ap.add_argument("--focal", type=int, default=64)
    ap.add_argument("--precond", default="jacobi", choices=["jacobi", "amg"])
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--device-resident", action="store_true",
                    help="cs_b200_solve_sources: RHS scattered on the device, node currents accumulated there")
    ap.add_argument("--check", type=int, default=0, help="verify this many columns against a grounded SciPy CG")
    args = ap.parse_args()
    import torch
    import circuitscape_b200 as cb
    from circuitscape_b200 import core, dist as D, graph

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    n = int(args.nodes)
    t0 = time.time()
    L = graph.power_law_laplacian(n, m=args.m, seed=11) if rank == 0 or world == 1 else None
    t_gen = time.time() - t0
    solver = cb.CUDASolver(device=local, precond=args.precond)
    t0 = time.time()
    comm = None
    if world > 1:
        import torch.distributed as dist
        dev = torch.device(f"cuda:{local}")
        dist.init_process_group("nccl", device_id=dev)

        def exchange(raw):
            t = torch.zeros(128, dtype=torch.uint8, device=dev)
            if raw is not None:
                t = torch.tensor(list(raw), dtype=torch.uint8, device=dev)
            dist.broadcast(t, src=0)
            return bytes(t.cpu().tolist())
        comm = D.Comm(local, rank, world, exchange)          # NCCL behind the C ABI (cs_b200_comm_*)
        meta = torch.zeros(2, dtype=torch.int64, device=dev)
        if rank == 0:
            meta = torch.tensor([L.shape[0], L.nnz], dtype=torch.int64, device=dev)
        dist.broadcast(meta, src=0)
        t0 = time.time()
        factor = comm.create_factor(L, solver, shape=(int(meta[0]), int(meta[1])))   # cs_b200_create_bcast
    else:
        factor = cb.construct_cholesky_factor(L, solver)
    t_setup = time.time() - t0
    focal = graph.focal_nodes(n, args.focal, seed=7)
    times = []
    for rep in range(args.reps + 1):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.time()
        if args.device_resident:
            factor.reset_currents()
        V, iters, relres, cols = core.all_to_one_batched(factor, focal, shard=(rank, world),
                                                         device_resident=args.device_resident,
                                                         accumulate=args.device_resident)
        if comm is not None and args.device_resident:
            comm.reduce_currents(factor)                     # end of the job: SUM / MAX of the current vectors
        torch.cuda.synchronize()
        dt = torch.tensor([time.time() - t0], dtype=torch.float64, device=f"cuda:{local}")
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        if rep:
            times.append(float(dt.item()))
    st = factor.stats()
    out = {"workload": f"power-law graph n={n} nnz={L.nnz if L is not None else None} all-to-one, {args.focal} focal nodes",
           "n_gpus": world, "precond": args.precond, "solves_per_s": args.focal / min(times),
           "s_per_pass_best": min(times), "s_per_pass_all": times, "iters_max": int(iters.max()),
           "iters_mean": float(iters.mean()), "relres_max": float(relres.max()),
           "gen_s": t_gen, "setup_s": t_setup, "kernel_ms_last_call": st.get("kernel_ms"),
           "through": ("cs_b200_solve_sources (device-resident, node currents accumulated on device)"
                       if args.device_resident else
                       "solve_linear_system n x k host batch (hook #2), includes H2D/D2H of n x k")}
    if args.check and rank == 0:
        import scipy.sparse.linalg as spla
        errs = []
        for j in range(min(args.check, len(cols))):
            f = focal[cols[j]]
            keep = np.ones(n, bool); keep[f] = False
            Lg = L[keep][:, keep].tocsr()
            b = np.zeros(n); b[focal] = 1.0
            dinv = 1.0 / Lg.diagonal()
            x, info = spla.cg(Lg, b[keep], rtol=1e-10, atol=0, maxiter=2000,
                              M=spla.LinearOperator(Lg.shape, lambda r: dinv * r))
            mine = V[j][focal != f] if args.device_resident else V[keep, j]
            ref = x[np.searchsorted(np.nonzero(keep)[0], focal[focal != f])] if args.device_resident else x
            errs.append(float(np.abs(mine - ref).max() / np.abs(x).max()))
        out["check_max_rel_err_vs_grounded_cpu_cg"] = max(errs)
    if rank == 0:
        print(json.dumps(out))
    factor.close()
    if world > 1:
        comm.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
