#!/usr/bin/env python
"""Large-configuration sanity + timing (BASELINE.json configs 3-5 shapes on ONE GPU):
  raster 4000x4000 (n = 1.6e7) fp64 and fp32, 16 pairs; power-law network, all-to-one."""
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
ap.add_argument("--nodes", type=int, default=2_000_000)
a = ap.parse_args()

if a.what == "raster":
    t = time.time()
    L, _ = graph.synthetic_raster_laplacian(a.rows, a.rows, seed=42)
    n = L.shape[0]
    print(f"assembled {a.rows}^2: n={n} nnz={L.nnz} in {time.time()-t:.1f}s", flush=True)
    nodes = graph.focal_nodes(n, 8, seed=7)
    src, dst = graph.all_pairs(nodes, limit=a.pairs)
    ref = None
    for prec in ("double", "single"):
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
else:
    # Barabasi-Albert-like power-law graph, m = 5, conductances U[0.1,1]
    rng = np.random.default_rng(11)
    n, m = a.nodes, 5
    t = time.time()
    tgt = np.empty(n * m, dtype=np.int64)
    # preferential attachment by sampling endpoints of existing edges (vectorised in chunks)
    srcs = np.repeat(np.arange(n), m)
    tgt[: m * (m + 1)] = rng.integers(0, m + 1, m * (m + 1))
    pos = m * (m + 1)
    chunk = 50_000
    while pos < n * m:
        hi = min(n * m, pos + chunk * m)
        pick = rng.integers(0, 2 * pos, hi - pos)
        ends = np.where(pick < pos, srcs[np.minimum(pick, pos - 1)], tgt[np.minimum(pick - pos, pos - 1)])
        tgt[pos:hi] = np.minimum(ends, srcs[pos:hi] - 1).clip(0)
        pos = hi
    keep = srcs != tgt
    W = sp.coo_matrix((rng.uniform(0.1, 1.0, keep.sum()), (srcs[keep], tgt[keep])), shape=(n, n)).tocsr()
    G = graph.laplacian(W + W.T)
    cc = graph.connected_components(G)
    big = max(cc, key=len) - 1
    G = G[big][:, big].tocsr()
    n = G.shape[0]
    deg = np.diff(G.indptr) - 1
    print(f"network: n={n} nnz={G.nnz} max degree {deg.max()} built in {time.time()-t:.1f}s", flush=True)
    focal = graph.focal_nodes(n, 8, seed=3)
    for prec_name in ("amg", "jacobi"):
        # all-to-one: ground focal[0] (Dirichlet row removed), +1 A at every other focal node
        gnd = focal[0]
        keepr = np.ones(n, dtype=bool); keepr[gnd] = False
        M = G[keepr][:, keepr].tocsr()
        b = np.zeros(n); b[focal[1:]] = 1.0
        b = b[keepr]
        t = time.time()
        with cb.B200Factor(M, cb.CUDASolver(precond=prec_name)) as f:
            ts = time.time() - t
            t = time.time()
            x, iters, relres = f.solve_rhs(b)
            dt = time.time() - t
            print(f"{prec_name}: setup {ts:.1f}s solve {dt*1e3:.1f} ms iters {iters.tolist()} relres {relres.max():.2e} "
                  f"true resid {np.linalg.norm(M @ x - b)/np.linalg.norm(b):.2e}", flush=True)
