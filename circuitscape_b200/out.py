"""Output stage of the path: the files the reference writes after a
solve, from the in-memory results of `core.solve` / `core.advanced_kernel` /
`core.onetoall_kernel`.

File names and contents follow src/out.jl:
  <pref>_resistances.out, <pref>_resistances_3columns.out      save_resistances   :454-465
  <pref>_curmap_<i>_<j>.asc, _voltmap_<i>_<j>.asc               write_grid         :321-386
  <pref>_cum_curmap.asc, _max_curmap.asc                        write_cum_maps     :467-482
  <pref>_node_currents_<i>_<j>.txt, _branch_currents_<i>_<j>.txt  write_currents   :115-124
  <pref>_voltages_<i>_<j>.txt                                   write_voltages     :412-419
with <pref> = output_file up to ".out".  Rasters are written as Arc/Info ASCII grids (the
reference goes through GDAL's AAIGrid driver; the header keys and NODATA = -9999 are the same,
number formatting is ours: repr-exact floats).  GeoTIFF output (`write_as_tif`) is not offered.
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np

NODATA = -9999.0


@dataclass
class RasterMeta:
    """src/io.jl RasterMeta, the fields an ASCII grid header needs."""
    ncols: int
    nrows: int
    xllcorner: float = 0.0
    yllcorner: float = 0.0
    cellsize: float = 1.0
    nodata: float = NODATA


def _pref(output_file):
    return output_file.split(".out")[0]          # split(cfg.output_file, ".out")[1]


def grid_filename(output_file, name="", voltage=False, cum=False, maxmap=False):
    """src/out.jl:325-336 -- cum wins over max wins over voltage, as in the reference."""
    s = "curmap"
    if cum:
        s = "cum_curmap"
    elif maxmap:
        s = "max_curmap"
    elif voltage:
        s = "voltmap"
    return f"{_pref(output_file)}_{s}{name}.asc"


def _fmt(x):
    return repr(float(x)) if x != int(x) or abs(x) >= 1e15 else str(int(x))


def write_asc(path, array, meta: RasterMeta):
    a = np.asarray(array, dtype=np.float64)
    if a.shape != (meta.nrows, meta.ncols):
        raise ValueError(f"array {a.shape} does not match the raster header {(meta.nrows, meta.ncols)}")
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "w") as f:
        f.write(f"ncols        {meta.ncols}\n")
        f.write(f"nrows        {meta.nrows}\n")
        f.write(f"xllcorner    {_fmt(meta.xllcorner)}\n")
        f.write(f"yllcorner    {_fmt(meta.yllcorner)}\n")
        f.write(f"cellsize     {_fmt(meta.cellsize)}\n")
        f.write(f"NODATA_value {_fmt(meta.nodata)}\n")
        for row in a:
            f.write(" ".join(_fmt(v) for v in row))
            f.write("\n")
    return path


def write_grid(cmap, name, output_file, meta, voltage=False, cum=False, maxmap=False):
    return write_asc(grid_filename(output_file, name, voltage, cum, maxmap), cmap, meta)


def compute_3col(r):
    from .core import compute_3col as _c3          # src/out.jl:12-26
    return _c3(np.asarray(r, dtype=np.float64))


def save_resistances(r, output_file):
    """src/out.jl:454-465 (space-delimited, like writedlm(f, r, ' '))."""
    pref = _pref(output_file)
    os.makedirs(os.path.dirname(os.path.abspath(pref)), exist_ok=True)
    paths = (f"{pref}_resistances.out", f"{pref}_resistances_3columns.out")
    for path, m in zip(paths, (np.asarray(r, dtype=np.float64), compute_3col(np.asarray(r, dtype=np.float64)))):
        with open(path, "w") as f:
            for row in m:
                f.write(" ".join(_fmt(v) for v in row))
                f.write("\n")
    return paths


def write_currents(node_curr_arr, branch_curr_arr, name, output_file):
    """src/out.jl:115-124: branch rows with |current| within 1e-6 of zero are dropped."""
    pref = _pref(output_file)
    os.makedirs(os.path.dirname(os.path.abspath(pref)), exist_ok=True)
    b = np.asarray(branch_curr_arr, dtype=np.float64).reshape(-1, 3)
    b = b[~np.isclose(b[:, 2], 0.0, atol=1e-6, rtol=0.0)]
    paths = (f"{pref}_node_currents{name}.txt", f"{pref}_branch_currents{name}.txt")
    for path, m in zip(paths, (np.asarray(node_curr_arr, dtype=np.float64).reshape(-1, 2), b)):
        with open(path, "w") as f:
            for row in m:
                f.write("\t".join(_fmt(v) for v in row))
                f.write("\n")
    return paths


def write_voltages(output_file, name, voltages, cc):
    """src/out.jl:412-419: (node id, voltage) rows of one component."""
    pref = _pref(output_file)
    os.makedirs(os.path.dirname(os.path.abspath(pref)), exist_ok=True)
    path = f"{pref}_voltages{name}.txt"
    with open(path, "w") as f:
        for node, v in zip(np.asarray(cc), np.asarray(voltages, dtype=np.float64)):
            f.write(f"{_fmt(node)}\t{_fmt(v)}\n")
    return path


def write_pairwise_outputs(result, output_file, meta: RasterMeta | None = None, write_cum=True, write_max=False):
    """Everything a `core.PairwiseOutput` holds, under the reference's file names.  Raster
    results need `meta`; network results (curmaps hold (nodes, currents) pairs) do not."""
    written = list(save_resistances(result.resistances, output_file))
    raster = meta is not None
    for (a, b), m in result.curmaps.items():
        if raster:
            written.append(write_grid(m, f"_{a}_{b}", output_file, meta))
        else:
            nodes, cur = m
            gr, gc, val = result.branch[(a, b)]
            written += write_currents(np.column_stack([nodes, cur]), np.column_stack([gr, gc, val]),
                                      f"_{a}_{b}", output_file)
    for (a, b), m in result.voltmaps.items():
        if raster:
            written.append(write_grid(m, f"_{a}_{b}", output_file, meta, voltage=True))
        else:
            nodes, v = m
            written.append(write_voltages(output_file, f"_{a}_{b}", v, nodes))
    if raster:
        if write_cum and result.cum_curmap is not None:
            written.append(write_grid(result.cum_curmap, "", output_file, meta, cum=True))
        if write_max and result.max_curmap is not None:
            written.append(write_grid(result.max_curmap, "", output_file, meta, maxmap=True))
    return written
