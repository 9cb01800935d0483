"""Why is cs_b200_create slower inside bench.py than in the C probe?  Same matrix, same library:
  python profiles/probes/py_setup_probe.py [--torch] [--rows 3163]
prints the wall time of two consecutive creates (cold / warm) with and without torch in the process."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
ap = argparse.ArgumentParser()
ap.add_argument("--torch", action="store_true")
ap.add_argument("--rows", type=int, default=3163)
a = ap.parse_args()
if a.torch:
    import torch
    torch.cuda.set_device(0)
    torch.zeros(1, device="cuda")
import circuitscape_b200 as cb  # noqa: E402
from circuitscape_b200 import graph  # noqa: E402

L, _ = graph.synthetic_raster_laplacian(a.rows, a.rows, seed=42)
print("modules:", "torch" in sys.modules, "threads:", os.cpu_count(), flush=True)
for rep in range(3):
    t0 = time.time()
    f = cb.construct_cholesky_factor(L, cb.CUDASolver())
    t1 = time.time()
    st = f.stats()
    f.close()
    print(f"rep {rep}: create {1e3 * (t1 - t0):.1f} ms wall, device-timed {st['setup_ms']:.1f} ms, close {1e3 * (time.time() - t1):.1f} ms", flush=True)
