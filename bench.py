#!/usr/bin/env python
"""bench.py -- pair-solves/sec of the focal-pair Laplacian solve loop on B200.

Contract (one JSON line on stdout from rank 0):
  python bench.py --gpus N --steps K --warmup W            (N>1 under torchrun)
  python bench.py --impl reference ...                     (CPU CG+AMG arm)

Workload (BASELINE.json configs[1], "C2"): 1000 x 1000 synthetic resistance raster
(R ~ U[1,10], seed 42), 8-neighbour average-conductance stencil, fp64, 5 focal nodes
(rng 7) -> 10 focal pairs per GPU.  A *step* = one pass of the hot path over that
batch: RHS build, batched PCG to rtol 1e-6, true-residual gate, resistance
extraction, node currents accumulated into the cumulative/max vectors -- all on the
device through `cs_b200_solve_pairs`.  The matrix/preconditioner is resident before
the timed region (the reference's "construct cholesky factor" is likewise once per
component, src/core.jl:379) and its cost is reported as setup_ms.
  value      pair-solves/s, whole job, device-timed (CUDA events on the solve stream)
  e2e        the same pairs through the plug-in hook `solve_linear_system(factor,
             matrix, rhs)` with HOST n x k RHS and solution buffers (H2D + D2H inside)
  roofline   dominant kernel (k_spmm) timed per launch with CUDA events in an
             instrumented repeat of the timed steps; plus the headline SpMV at
             10^7 nodes (3163^2) under `spmv_1e7`
  cpu_baseline  the oracle's CG+AMG port on the host cores, bounded sample
Multi-GPU: pairs are sharded over ranks (10 per GPU, weak scaling), the CSR is
broadcast once over NCCL, resistances are all-gathered and the cumulative / max
current vectors all-reduced inside every step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

L2_BYTES = 126e6


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=1000)
    ap.add_argument("--cols", type=int, default=1000)
    ap.add_argument("--pairs-per-gpu", type=int, default=10)
    ap.add_argument("--precision", default="double")
    ap.add_argument("--precond", default="amg", choices=["amg", "jacobi"])
    ap.add_argument("--rtol", type=float, default=1e-6)
    ap.add_argument("--loop", default="device", choices=["device", "chunk", "plain"],
                    help="PCG loop control: device-side WHILE graph | host-polled graph chunks | plain launches")
    ap.add_argument("--skip-spmv1e7", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=0, help="pairs in the CPU sample (0 = #cores)")
    ap.add_argument("--cpu-direct", action="store_true",
                    help="also time the CHOLMOD-like CPU path (factor once + batched solves, src/core.jl:379,448-463) "
                         "with SciPy SuperLU; ~15 s and ~3 GB at 1000^2, off by default")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        return json.load(open(p))["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def b_spmm(n, nnz, k, sv):
    """SURVEY.md §8d / BASELINE.md: algorithmic bytes of one SpMM launch."""
    return nnz * (sv + 4) + (n + 1) * 4 + 2 * n * k * sv


def workload(args, total_pairs):
    from circuitscape_b200 import graph
    L, _ = graph.synthetic_raster_laplacian(args.rows, args.cols, seed=42,
                                            dtype=np.float64 if args.precision == "double" else np.float32)
    npts = 2
    while npts * (npts - 1) // 2 < total_pairs:
        npts += 1
    nodes = graph.focal_nodes(L.shape[0], npts, seed=7)
    src, dst = graph.all_pairs(nodes, limit=total_pairs)
    return L, src, dst


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.proc = None
        self.lines = []
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------
# CPU arm: the oracle's CG+AMG port, one pair per core (mirrors src/core.jl:262-272)
# ---------------------------------------------------------------------------
_CPU = {}


def _cpu_init():
    try:                                   # one BLAS thread per worker: no oversubscription
        from threadpoolctl import threadpool_limits
        _CPU["_tp"] = threadpool_limits(1)
    except Exception:
        pass


def _cpu_one(i):
    from oracle import amg
    A, ml, src, dst = _CPU["A"], _CPU["ml"], _CPU["src"], _CPU["dst"]
    n = A.shape[0]
    b = np.zeros(n); b[src[i]] = -1.0; b[dst[i]] = 1.0
    v, it = amg.pcg(A, b, ml, rtol=1e-6, itmax=100_000)
    res = np.linalg.norm(A @ v - b) / np.sqrt(2.0)
    assert res < 1e-4                                    # src/core.jl:640-641
    return float(v[dst[i]] - v[src[i]]), it


def cpu_cg_amg(L, src, dst, sample, repeats=1):
    """returns dict(value pairs/s, cores, setup_s, iters, R).  Setup (AMG hierarchy,
    src/core.jl:164-167 "construct preconditioner") is excluded like the GPU setup."""
    import multiprocessing as mp
    from oracle import amg
    cores = len(os.sched_getaffinity(0))
    A = L.astype(np.float64).tocsr().copy()
    A.data = A.data + np.finfo(np.float64).eps * np.linalg.norm(A.data)     # src/core.jl:161
    t0 = time.time()
    ml = amg.smoothed_aggregation(A)
    setup = time.time() - t0
    sample = min(sample, len(src))
    _CPU.update(A=A, ml=ml, src=src, dst=dst)
    ctx = mp.get_context("fork")
    times, out = [], None
    with ctx.Pool(min(cores, sample), initializer=_cpu_init) as pool:
        pool.map(_cpu_one, range(min(cores, sample)))      # warm the workers' caches / page-in
        for _ in range(repeats):
            t0 = time.time()
            out = pool.map(_cpu_one, range(sample), chunksize=1)
            times.append(time.time() - t0)
    wall = float(np.mean(times))
    return dict(value=sample / wall, cores=min(cores, sample), host_cores=cores, setup_s=setup, wall_s=wall,
                times=times, iters=[o[1] for o in out], R=[o[0] for o in out], sample=sample,
                levels=[l.n for l in ml.levels])


def run_reference(args):
    """--impl reference: the reference's CPU CG+AMG path (oracle port; Julia is not in
    the image) on the same workload, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    L, src, dst = workload(args, args.pairs_per_gpu * args.gpus)
    sample = args.cpu_sample or min(len(src), len(os.sched_getaffinity(0)))
    r = cpu_cg_amg(L, src, dst, sample, repeats=max(1, args.steps))
    ms = r["wall_s"] * 1e3
    line = {
        "impl": "reference", "metric": "pair_solves_per_sec", "value": r["value"], "unit": "pair-solves/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args, L, len(src)),
        "cpu_baseline": {"value": r["value"], "unit": "pair-solves/s", "cores": r["cores"], "kind": "port",
                         "sample": f"{r['sample']} of {len(src)} pairs per step, one pair per process "
                                   f"(SA-AMG+symmetric-GS PCG rtol 1e-6; AMG setup {r['setup_s']:.1f}s excluded; "
                                   f"iterations {r['iters']})"},
        "e2e": {"value": r["value"], "unit": "pair-solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def config_dict(args, L, npairs):
    return {"workload": f"C2: {args.rows}x{args.cols} synthetic raster (R~U[1,10] seed 42), 8-neighbour "
                        f"avg-conductance, {args.pairs_per_gpu} focal pairs per GPU, {args.precision}",
            "n": int(L.shape[0]), "nnz": int(L.nnz), "pairs_total": int(npairs), "rtol": args.rtol,
            "preconditioner": args.precond, "parallelism": f"pair-shard x{args.gpus}",
            "l2_policy": "working set per iteration (matrix + 4 panels, ~0.4 GB) exceeds the 126 MB L2"}


# ---------------------------------------------------------------------------
def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import circuitscape_b200 as cb
    from circuitscape_b200 import dist as cdist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    distributed = world > 1
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback on the product path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if distributed:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    else:
        dist = None

    total_pairs = args.pairs_per_gpu * world
    solver = cb.CUDASolver(precision=args.precision, device=local, rtol=args.rtol, precond=args.precond,
                           use_graph={"device": True, "chunk": "chunk", "plain": False}[args.loop])
    t_asm = time.time()
    L = src = dst = None
    if rank == 0:
        L, src, dst = workload(args, total_pairs)
    t_asm = time.time() - t_asm
    # ---- replicate the operator: one NCCL broadcast of the CSR -------
    t0 = time.time()
    if distributed:
        n, nnz, rp, ci, va = cdist.broadcast_csr(L, dist, dev)
        pairs = torch.zeros((2, total_pairs), dtype=torch.int64, device=dev)
        if rank == 0:
            pairs = torch.as_tensor(np.stack([src, dst]), device=dev)
        dist.broadcast(pairs, src=0)
        src, dst = pairs[0].cpu().numpy(), pairs[1].cpu().numpy()
        factor = cdist.factor_from_device(n, nnz, rp, ci, va, solver)
    else:
        n, nnz = L.shape[0], L.nnz
        factor = cb.construct_cholesky_factor(L, solver)
    torch.cuda.synchronize()
    setup_s = time.time() - t0
    mine = cdist.shard_pairs(total_pairs, rank, world)
    msrc, mdst = src[mine], dst[mine]
    ext = torch.cuda.ExternalStream(factor.stream_ptr(), device=dev)

    def step():
        factor.reset_currents()
        out = factor.solve_pairs(msrc, mdst, accumulate=True)
        if distributed:
            R = cdist.gather_pairs(mine, out["R"], total_pairs, dist, device=dev)
            cdist.reduce_currents(factor, dist)
        else:
            R = out["R"]
        return R, out, factor.stats()

    def timed(fn, steps):
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        t0 = time.time()
        res = [fn() for _ in range(steps)]
        e1.record(ext)
        torch.cuda.synchronize()
        if distributed:
            dist.barrier()
        wall = time.time() - t0
        ms = e0.elapsed_time(e1)
        if distributed:
            t = torch.tensor([ms, wall * 1e3], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, wall = float(t[0]), float(t[1]) / 1e3
        return ms, wall, res

    # clocks are sampled from the first warm-up step to the end of the timed region (the
    # timed region alone can be shorter than nvidia-smi's start-up + sampling period)
    sampler = ClockSampler(local) if rank == 0 else None
    for _ in range(args.warmup):
        step()
    t_w = time.time()
    while rank == 0 and sampler and len(sampler.lines) < 2 and time.time() - t_w < 3.0:
        factor.solve_pairs(msrc, mdst)        # local work only (no collective): load until samples arrive
    ms, wall, res = timed(step, args.steps)
    clocks = sampler.stop() if sampler else None
    R, out, st = res[-1]
    launches = sum(r[2]["kernel_launches"] for r in res)
    iters = int(out["iters"].sum())
    value = total_pairs * args.steps / (ms / 1e3)

    # ---- e2e: plug-in hook #2 with host RHS / solution buffers --------------------
    k = len(msrc)
    tdt = torch.float64 if factor.dtype == np.float64 else torch.float32
    # pinned host buffers, column-major n x k (a Julia Matrix): torch (k, n) row-major == (n, k) F-order
    rhs = torch.zeros((k, n), dtype=tdt).pin_memory().numpy().T
    lhs_buf = torch.zeros((k, n), dtype=tdt).pin_memory().numpy().T
    rhs[msrc, np.arange(k)] = -1.0
    rhs[mdst, np.arange(k)] = 1.0

    def step_e2e():
        lhs, _, _ = factor.solve_rhs(rhs, out=lhs_buf)          # hook #2 on pinned host buffers
        r = lhs[mdst, np.arange(k)] - lhs[msrc, np.arange(k)]   # src/core.jl:466-472, 486-492
        if distributed:
            r = cdist.gather_pairs(mine, r, total_pairs, dist, device=dev)
        return r, factor.stats()

    step_e2e()
    ms_e, wall_e, res_e = timed(step_e2e, max(1, min(args.steps, 3)))
    nst_e = max(1, min(args.steps, 3))
    e2e_value = total_pairs * nst_e / (wall_e if wall_e * 1e3 > ms_e else ms_e / 1e3)
    e2e = {"value": e2e_value, "unit": "pair-solves/s",
           "h2d_bytes_per_step": int(res_e[-1][1]["h2d_bytes"]) * world,
           "d2h_bytes_per_step": int(res_e[-1][1]["d2h_bytes"]) * world,
           "ms_per_step": wall_e * 1e3 / nst_e,
           "through": "solve_linear_system(factor, matrix, rhs::Matrix) with pinned host n x k buffers"}
    assert np.abs(np.asarray(res_e[-1][0]) - np.asarray(R)).max() <= 1e-6 * np.abs(R).max()

    # ---- roofline of the dominant kernel: instrumented repeat of one timed step ---
    peak, peak_src = peaks()
    roof = None
    extra = {}
    if rank == 0:
        factor.profile_spmm(True)
        factor.reset_currents()
        factor.solve_pairs(msrc, mdst, accumulate=True)     # rank-local repeat of the step's solve
        pbytes = factor.profile_bytes()
        pms, pl = factor.profile_spmm(False)
        sv = 8 if args.precision == "double" else 4
        widths = []
        rem = k
        while rem > 0:
            w = 8
            while w > rem:
                w //= 2
            widths.append(w); rem -= w
        avg_bytes = pbytes / max(pl, 1)
        achieved = pbytes / (pms * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r1_traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath)).get("c2_finest_level_spmm")
            if tj:
                traffic = tj.get("traffic_bytes_per_launch")
        roof = {"bound": "hbm",
                "kernel": "k_spmm_win on the finest-level operator, every epilogue of the AMG-PCG iteration "
                          "(fp64 CG / residual gate, fp32 residual + Jacobi sweep of the V-cycle), panels "
                          + "+".join(map(str, widths)),
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "launches": int(pl),
                "avg_launch_ms": pms / max(pl, 1), "algorithmic_bytes_per_launch": avg_bytes,
                "spmm_share_of_step": pms / (ms / args.steps),
                "note": "per-launch CUDA events on the solve stream in an instrumented repeat of the timed "
                        "step (plain launches, same kernels as the graph); bytes = nnz(s_v+4)+(n+1)4+panel "
                        "passes summed per launch by the library; traffic = ncu dram bytes per launch averaged "
                        "over the same kernels (profiles/r1_traffic.json); 1000^2 operands are partly "
                        "L2-resident (see spmv_1e7 for the HBM-bound size)"}
        # supplementary: the same pairs through the superposition driver (one solve per focal
        # NODE; not the headline -- the headline counts one linear solve per pair, as the reference does)
        try:
            nn_, inv_ = np.unique(np.concatenate([msrc, mdst]), return_inverse=True)
            pi_, pj_ = inv_[:len(msrc)], inv_[len(msrc):]
            factor.reset_currents()
            osup = factor.solve_pairs_superposed(nn_, pi_, pj_, accumulate=True)
            tsup = []
            for _ in range(3):
                factor.reset_currents()
                t0 = time.time()
                osup = factor.solve_pairs_superposed(nn_, pi_, pj_, accumulate=True)
                tsup.append(time.time() - t0)
            extra["superposed_driver"] = {
                "pair_solves_per_s": len(msrc) / min(tsup), "ms_per_step": min(tsup) * 1e3,
                "point_solves": int(len(nn_) - 1), "pairs": int(len(msrc)),
                "max_rel_dev_of_R": float(np.abs(osup["R"] - out["R"]).max() / np.abs(out["R"]).max()),
                "relres_max": float(osup["relres"].max()),
                "note": "host wall clock around cs_b200_solve_pairs_superposed on rank 0's pairs"}
        except Exception as exc:                                   # never let a side leg break the line
            extra["superposed_driver"] = {"error": repr(exc)}
        for kk in (1, 8):
            t_it = factor.bench_cg_iter(kk, reps=50)
            b_it = b_spmm(n, nnz, kk, sv) + 8 * n * kk * sv + 2 * n * sv
            extra[f"cg_iter_k{kk}"] = {"ms": t_it, "GB/s": b_it / (t_it * 1e-3) / 1e9,
                                       "algorithmic_bytes": b_it}

    # ---- headline SpMV at 10^7 nodes ---------------------------------------------
    spmv = None
    if rank == 0 and not args.skip_spmv1e7:
        factor.close()
        from circuitscape_b200 import graph
        t0 = time.time()
        L7, _ = graph.synthetic_raster_laplacian(3163, 3163, seed=42)
        with cb.construct_cholesky_factor(L7, cb.CUDASolver(device=local, precond="jacobi")) as f7:
            n7, nnz7 = L7.shape[0], L7.nnz
            spmv = {"n": n7, "nnz": nnz7, "assemble_upload_s": time.time() - t0, "peak": peak, "peak_source": peak_src}
            for kk in (1, 8):
                t = f7.bench_spmm(kk, reps=20, flush_l2=True)
                b = b_spmm(n7, nnz7, kk, 8)
                spmv[f"k{kk}"] = {"ms": t, "algorithmic_bytes": b, "GB/s": b / (t * 1e-3) / 1e9,
                                  "frac": b / (t * 1e-3) / 1e9 / peak}
            for kk in (1, 8):
                t = f7.bench_cg_iter(kk, reps=20)
                b = b_spmm(n7, nnz7, kk, 8) + 8 * n7 * kk * 8 + 2 * n7 * 8
                spmv[f"cg_iter_k{kk}"] = {"ms": t, "algorithmic_bytes": b, "GB/s": b / (t * 1e-3) / 1e9,
                                          "frac": b / (t * 1e-3) / 1e9 / peak}
        del L7

    # ---- CPU baseline (rank 0, N = 1 only) ----------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        sample = args.cpu_sample or min(len(src), len(os.sched_getaffinity(0)))
        r = cpu_cg_amg(L, src, dst, sample)
        cpu = {"value": r["value"], "unit": "pair-solves/s", "cores": r["cores"], "kind": "port",
               "sample": f"{r['sample']} of {len(src)} pairs, one pair per process on {r['host_cores']} host cores, "
                         f"SA-AMG(sym. GS, pinv coarse)-PCG rtol 1e-6 (oracle/amg.py; Julia absent); "
                         f"AMG setup {r['setup_s']:.1f}s excluded; {r['wall_s']:.1f}s wall; iterations {r['iters']}",
               "max_rel_dev_from_gpu_R": float(np.max(np.abs(np.array(r["R"]) - np.asarray(R)[:r["sample"]])
                                                      / np.asarray(R)[:r["sample"]]))}

    if rank == 0 and world == 1 and args.cpu_direct:
        import scipy.sparse as _sp
        import scipy.sparse.linalg as _spla
        t0 = time.time()
        Md = (L.astype(np.float64) + 10 * np.finfo(np.float64).eps * _sp.identity(n)).tocsc()   # core.jl:521
        lu = _spla.splu(Md, permc_spec="MMD_AT_PLUS_A", diag_pivot_thresh=0.0, options=dict(SymmetricMode=True))
        tf = time.time() - t0
        rhs_d = np.zeros((n, len(src)))
        rhs_d[src, np.arange(len(src))] = -1.0
        rhs_d[dst, np.arange(len(src))] = 1.0
        t0 = time.time()
        Xd = lu.solve(rhs_d)
        tsv = time.time() - t0
        Rd = Xd[dst, np.arange(len(src))] - Xd[src, np.arange(len(src))]
        extra["cpu_direct"] = {"kind": "port (SciPy SuperLU standing in for CHOLMOD)", "cores": 1,
                               "factor_s": tf, "solve_s": tsv, "pairs": len(src),
                               "pair_solves_per_s_incl_factor": len(src) / (tf + tsv),
                               "pair_solves_per_s_excl_factor": len(src) / tsv,
                               "max_rel_dev_from_gpu_R": float(np.max(np.abs(Rd - np.asarray(R)[:len(src)]) / Rd))}
        del lu, Xd, rhs_d, Md
    if rank == 0:
        line = {
            "metric": "pair_solves_per_sec", "value": value, "unit": "pair-solves/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64" if args.precision == "double" else "f32", "data": "synthetic",
            "config": config_dict(args, L, total_pairs), "clocks": clocks, "e2e": e2e,
            "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu, "spmv_1e7": spmv,
            "detail": {"iterations_per_step_rank0": iters, "setup_s": setup_s, "assemble_s": t_asm,
                       "wall_s_timed_region": wall, "setup_ms_device": st["setup_ms"],
                       "relres_max": float(out["relres"].max()), "R_first": [float(x) for x in np.asarray(R)[:3]],
                       **extra},
        }
        print(json.dumps(line), flush=True)
    if distributed:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
