"""C2 workload (1000^2, 5 focal nodes -> 10 pairs) through the per-pair driver and through the
superposition driver (cs_b200_solve_pairs_superposed): device time per pass, agreement of R and
of the cumulative current map."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import circuitscape_b200 as cb
from circuitscape_b200 import graph

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
npts = int(sys.argv[2]) if len(sys.argv) > 2 else 5
L, _ = graph.synthetic_raster_laplacian(rows, rows, seed=42)
nodes = graph.focal_nodes(L.shape[0], npts, seed=7)
src, dst = graph.all_pairs(nodes)
nn, inv = np.unique(np.concatenate([src, dst]), return_inverse=True)
pi, pj = inv[:len(src)], inv[len(src):]
with cb.B200Factor(L, cb.CUDASolver()) as f:
    for name, call in (("per-pair", lambda: f.solve_pairs(src, dst, accumulate=True)),
                       ("superposed", lambda: f.solve_pairs_superposed(nn, pi, pj, accumulate=True))):
        f.reset_currents(); o = call()
        ts = []
        for _ in range(5):
            f.reset_currents(); t = time.time(); o = call(); ts.append(time.time() - t)
        cum, _ = f.read_currents()
        st = f.stats()
        print(f"{name:10s}: {len(src)} pairs in {min(ts)*1e3:.2f} ms = {len(src)/min(ts):.0f} pair-solves/s "
              f"(kernel {st['kernel_ms']:.2f} ms, launches {st['kernel_launches']}); relres max {o['relres'].max():.2e}", flush=True)
        if name == "per-pair":
            R0, c0 = o["R"].copy(), cum.copy()
        else:
            print(f"   max rel dev of R {np.abs(o['R']-R0).max()/R0.max():.2e}, of the cumulative map {np.abs(cum-c0).max()/c0.max():.2e}")
