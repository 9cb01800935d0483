#!/usr/bin/env python
"""AMG solve at scale under different window masks / panel widths (debug aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import circuitscape_b200 as cb
from circuitscape_b200 import graph

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
masks = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 1, 2, 4, 8, 15]
L, _ = graph.synthetic_raster_laplacian(rows, rows, seed=42)
n = L.shape[0]
nodes = graph.focal_nodes(n, 5, seed=7)
src, dst = graph.all_pairs(nodes)
for mask in masks:
    os.environ["CS_B200_WIN_MASK"] = str(mask)
    for pw in (8, 4):
        t = time.time()
        with cb.B200Factor(L, cb.CUDASolver(precond="amg", panel_width=pw, itmax=300)) as f:
            o = f.solve_pairs(src[:pw], dst[:pw], raise_on_residual=False)
        print(f"mask={mask:2d} pw={pw}: iters {o['iters'].tolist()} relres max {o['relres'].max():.2e} R0 {o['R'][0]:.9f}  ({time.time()-t:.1f}s)", flush=True)
