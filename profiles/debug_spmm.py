#!/usr/bin/env python
"""SpMM parity at full size: windowed (TMA) vs plain kernel vs scipy, all panel widths."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import circuitscape_b200 as cb
from circuitscape_b200 import graph

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
L, _ = graph.synthetic_raster_laplacian(rows, rows, seed=42)
n = L.shape[0]
rng = np.random.default_rng(0)
for window in ("on", "off"):
    with cb.B200Factor(L, cb.CUDASolver(window=window)) as f:
        for k in (1, 2, 4, 8):
            X = rng.standard_normal((n, k))
            Y = f.spmm(X)
            ref = L @ X
            err = np.abs(Y - ref)
            bad = np.argwhere(err > 1e-9)
            print(f"window={window} k={k}: max err {err.max():.3e}  bad entries {len(bad)}"
                  + (f" first bad row {bad[0]} rows%256 {sorted(set((bad[:,0] % 256).tolist()))[:10]} blocks {sorted(set((bad[:,0]//256).tolist()))[:10]}" if len(bad) else ""))
