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
number formatting is ours: repr-exact floats).  With `write_as_tif` the same arrays go to single-band
GeoTIFFs (src/out.jl:338,378,483-531: GTiff driver, nodata -9999, the input's geotransform and WKT);
GDAL is not in the image, so `write_tif` emits the baseline TIFF + GeoTIFF tags itself (deflate or no
compression where the reference asks GDAL for LZW -- all lossless, the decoded pixels are the same).
"""
from __future__ import annotations

import os
import struct
import zlib
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
    transform: tuple | None = None   # GDAL geotransform (x0, dx, rx, y0, ry, dy); None: from the corner + cellsize
    wkt: str = ""                    # projection of the input raster, carried to GeoTIFF outputs verbatim

    def geotransform(self):
        if self.transform is not None:
            t = tuple(float(v) for v in self.transform)
            if len(t) != 6:
                raise ValueError("a geotransform has 6 coefficients")
            return t
        return (float(self.xllcorner), float(self.cellsize), 0.0,
                float(self.yllcorner) + self.nrows * float(self.cellsize), 0.0, -float(self.cellsize))


def _pref(output_file):
    return output_file.split(".out")[0]          # split(cfg.output_file, ".out")[1]


def grid_filename(output_file, name="", voltage=False, cum=False, maxmap=False, tif=False):
    """src/out.jl:325-336 -- cum wins over max wins over voltage, as in the reference."""
    s = "curmap"
    if cum:
        s = "cum_curmap"
    elif maxmap:
        s = "max_curmap"
    elif voltage:
        s = "voltmap"
    return f"{_pref(output_file)}_{s}{name}{'.tif' if tif else '.asc'}"


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


# TIFF field types
_T_SHORT, _T_LONG, _T_DOUBLE, _T_ASCII = 3, 4, 12, 2


def write_tif(path, array, meta: RasterMeta, compress="deflate", rows_per_strip=None):
    """Single-band GeoTIFF of `array` (float32 stays float32, everything else float64): little-endian
    baseline TIFF, one strip per `rows_per_strip` rows, SampleFormat = IEEE float, GDAL_NODATA = the
    header's nodata, ModelPixelScale + ModelTiepoint (or ModelTransformation for a rotated geotransform),
    and -- when the header carries a WKT -- a user-defined projected CRS whose PCSCitationGeoKey holds
    `ESRI PE String = <wkt>`, the form GDAL turns back into the same WKT.  Mirrors src/out.jl:483-531."""
    a = np.asarray(array)
    a = np.ascontiguousarray(a, dtype="<f4" if a.dtype == np.float32 else "<f8")
    if a.shape != (meta.nrows, meta.ncols):
        raise ValueError(f"array {a.shape} does not match the raster header {(meta.nrows, meta.ncols)}")
    if compress not in ("deflate", "none"):
        raise ValueError("compress must be 'deflate' or 'none'")
    nrows, ncols = a.shape
    bps = a.dtype.itemsize * 8
    if rows_per_strip is None:
        rows_per_strip = max(1, min(nrows, (1 << 20) // max(1, ncols * a.dtype.itemsize)))
    strips = []
    for r0 in range(0, nrows, rows_per_strip):
        raw = a[r0:r0 + rows_per_strip].tobytes()
        strips.append(zlib.compress(raw, 6) if compress == "deflate" else raw)
    x0, dx, rx, y0, ry, dy = meta.geotransform()
    # (tag, type, values); ASCII values are bytes incl. the terminating NUL
    tags = [(256, _T_LONG, [ncols]), (257, _T_LONG, [nrows]), (258, _T_SHORT, [bps]),
            (259, _T_SHORT, [8 if compress == "deflate" else 1]), (262, _T_SHORT, [1]),
            (273, _T_LONG, None),                      # strip offsets: patched below
            (277, _T_SHORT, [1]), (278, _T_LONG, [rows_per_strip]),
            (279, _T_LONG, [len(b) for b in strips]), (284, _T_SHORT, [1]), (339, _T_SHORT, [3])]
    if rx == 0.0 and ry == 0.0:
        tags.append((33550, _T_DOUBLE, [dx, -dy, 0.0]))
        tags.append((33922, _T_DOUBLE, [0.0, 0.0, 0.0, x0, y0, 0.0]))
    else:
        tags.append((34264, _T_DOUBLE, [dx, rx, 0.0, x0, ry, dy, 0.0, y0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0]))
    if meta.wkt:
        cit = ("ESRI PE String = " + meta.wkt + "|").encode("ascii", "replace") + b"\0"
        # GeoKeyDirectory: version 1.1.0, 4 keys: model = projected, raster = PixelIsArea,
        # ProjectedCSType = user-defined, PCSCitation -> GeoAsciiParams[0 : len - 1]
        keys = [1, 1, 0, 4, 1024, 0, 1, 1, 1025, 0, 1, 1, 3072, 0, 1, 32767, 3073, 34737, len(cit) - 1, 0]
        tags.append((34735, _T_SHORT, keys))
        tags.append((34737, _T_ASCII, cit))
    else:
        tags.append((34735, _T_SHORT, [1, 1, 0, 1, 1025, 0, 1, 1]))
    tags.append((42113, _T_ASCII, _fmt(meta.nodata).encode("ascii") + b"\0"))
    tags.sort(key=lambda t: t[0])

    fmt = {_T_SHORT: "H", _T_LONG: "I", _T_DOUBLE: "d"}
    size = {_T_SHORT: 2, _T_LONG: 4, _T_DOUBLE: 8, _T_ASCII: 1}
    nstrips = len(strips)
    ifd_off = 8
    ifd_len = 2 + 12 * len(tags) + 4
    # layout: header | IFD | out-of-line tag values | strips
    extra_off = ifd_off + ifd_len
    extra = bytearray()
    entries = []
    placed = {}
    for tag, typ, vals in tags:
        count = nstrips if tag == 273 else len(vals)
        nbytes = count * size[typ]
        if nbytes <= 4:
            placed[tag] = None
        else:
            if len(extra) % 2:
                extra += b"\0"
            placed[tag] = extra_off + len(extra)
            extra += b"\0" * nbytes
        entries.append((tag, typ, count))
    if len(extra) % 2:
        extra += b"\0"
    data_off = extra_off + len(extra)
    offsets, o = [], data_off
    for b in strips:
        offsets.append(o)
        o += len(b) + (len(b) % 2)
    if o >= 1 << 32:
        raise ValueError("raster too large for a classic (32-bit offset) TIFF")

    def payload(tag, typ, vals):
        if tag == 273:
            vals = offsets
        if typ == _T_ASCII:
            return bytes(vals)
        return struct.pack("<" + fmt[typ] * len(vals), *vals)

    ifd = bytearray(struct.pack("<H", len(tags)))
    for (tag, typ, vals), (_, _, count) in zip(tags, entries):
        body = payload(tag, typ, vals)
        if placed[tag] is None:
            ifd += struct.pack("<HHI", tag, typ, count) + body.ljust(4, b"\0")
        else:
            ifd += struct.pack("<HHII", tag, typ, count, placed[tag])
            at = placed[tag] - extra_off
            extra[at:at + len(body)] = body
    ifd += struct.pack("<I", 0)
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "wb") as f:
        f.write(b"II" + struct.pack("<HI", 42, ifd_off))
        f.write(ifd)
        f.write(extra)
        for b in strips:
            f.write(b)
            if len(b) % 2:
                f.write(b"\0")
    return path


def write_grid(cmap, name, output_file, meta, voltage=False, cum=False, maxmap=False, write_as_tif=False):
    """src/out.jl:321-345: one raster under the reference's name, `.asc` or (cfg.write_as_tif) `.tif`."""
    fn = grid_filename(output_file, name, voltage, cum, maxmap, tif=write_as_tif)
    return write_tif(fn, cmap, meta) if write_as_tif else write_asc(fn, cmap, meta)


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


def write_pairwise_outputs(result, output_file, meta: RasterMeta | None = None, write_cum=True, write_max=False,
                           write_as_tif=False):
    """Everything a `core.PairwiseOutput` holds, under the reference's file names.  Raster
    results need `meta`; network results (curmaps hold (nodes, currents) pairs) do not."""
    written = list(save_resistances(result.resistances, output_file))
    raster = meta is not None
    for (a, b), m in result.curmaps.items():
        if raster:
            written.append(write_grid(m, f"_{a}_{b}", output_file, meta, write_as_tif=write_as_tif))
        else:
            nodes, cur = m
            gr, gc, val = result.branch[(a, b)]
            written += write_currents(np.column_stack([nodes, cur]), np.column_stack([gr, gc, val]),
                                      f"_{a}_{b}", output_file)
    for (a, b), m in result.voltmaps.items():
        if raster:
            written.append(write_grid(m, f"_{a}_{b}", output_file, meta, voltage=True, write_as_tif=write_as_tif))
        else:
            nodes, v = m
            written.append(write_voltages(output_file, f"_{a}_{b}", v, nodes))
    if raster:
        if write_cum and result.cum_curmap is not None:
            written.append(write_grid(result.cum_curmap, "", output_file, meta, cum=True, write_as_tif=write_as_tif))
        if write_max and result.max_curmap is not None:
            written.append(write_grid(result.max_curmap, "", output_file, meta, maxmap=True, write_as_tif=write_as_tif))
    return written
