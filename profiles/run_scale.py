#!/usr/bin/env python
"""Large-configuration sanity + timing (BASELINE.json configs 3-5 shapes on ONE GPU):
  raster 4000x4000 (n = 1.6e7), 16 pairs (config C5, the network, is profiles/run_network.py)."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.sparse as sp
import circuitscape_b200 as cb
from circuitscape_b200 import graph

ap = argparse.ArgumentParser()
ap.add_argument("--what", default="raster", choices=["raster", "network"])
ap.add_argument("--rows", type=int, default=4000)
ap.add_argument("--pairs", type=int, default=16)
ap.add_argument("--device-assembly", action="store_true",
                help="also build the handle with cs_b200_create_from_raster (Laplacian assembled on the GPU) and compare")
ap.add_argument("--precisions", default="double", help="comma list of double,single (single = fp32 host buffers, fp64 on device)")
a = ap.parse_args()

if a.what == "raster":
    t = time.time()
    L, g = graph.synthetic_raster_laplacian(a.rows, a.rows, seed=42)
    n = L.shape[0]
    print(f"assembled {a.rows}^2: n={n} nnz={L.nnz} in {time.time()-t:.1f}s", flush=True)
    nodes = graph.focal_nodes(n, 8, seed=7)
    src, dst = graph.all_pairs(nodes, limit=a.pairs)
    ref = None
    for prec in a.precisions.split(","):
        t = time.time()
        with cb.B200Factor(L.astype(np.float64 if prec == "double" else np.float32), cb.CUDASolver(precision=prec)) as f:
            ts = time.time() - t
            o = f.solve_pairs(src, dst, accumulate=True)      # warm
            t = time.time()
            o = f.solve_pairs(src, dst, accumulate=True)
            dt = time.time() - t
            st = f.stats()
            R = o["R"].astype(np.float64)
            if ref is None:
                ref = R
            print(f"{prec}: setup {ts:.1f}s  {a.pairs} pairs in {dt*1e3:.1f} ms = {a.pairs/dt:.1f} pair-solves/s; "
                  f"iters {o['iters'].tolist()} relres max {o['relres'].max():.2e}; "
                  f"max rel dev of R from fp64 {np.abs(R-ref).max()/ref.max():.2e}; launches {st['kernel_launches']}", flush=True)
            for k in (1, 8):
                print(f"   spmm k={k}: {f.bench_spmm(k, reps=10, flush_l2=True):.3f} ms   cg_iter: {f.bench_cg_iter(k, reps=10):.3f} ms", flush=True)
    if a.device_assembly:
        t = time.time()
        with cb.B200Factor.from_raster(g, cb.CUDASolver()) as f:
            ts = time.time() - t
            o = f.solve_pairs(src, dst, accumulate=True)
            print(f"device assembly: handle from the {a.rows}^2 conductance raster in {ts:.1f}s (n={f.n}); "
                  f"max rel dev of R from the host-assembled run {np.abs(o['R'] - ref).max() / ref.max():.2e}", flush=True)
else:
    raise SystemExit("network configuration: use profiles/run_network.py")
