"""which setup phase makes the SECOND create of a process slow?  (bench.py: create_s > create_first)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import circuitscape_b200 as cb
from circuitscape_b200 import graph
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
L, _ = graph.synthetic_raster_laplacian(rows, rows, seed=42)
for rep in range(3):
    if rep >= 1:
        os.environ["CS_B200_VERBOSE"] = "2"
    t0 = time.time()
    f = cb.construct_cholesky_factor(L, cb.CUDASolver())
    t1 = time.time()
    st = f.stats()
    print(f"rep {rep}: create {1e3 * (t1 - t0):.1f} ms wall, device-timed {st['setup_ms']:.1f} ms", flush=True)
    t2 = time.time()
    f.close()
    print(f"rep {rep}: close {1e3 * (time.time() - t2):.1f} ms", flush=True)
