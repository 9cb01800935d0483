#!/usr/bin/env python
"""Pack the reference's own golden test vectors for the hot path into one .npz.

TEST INFRASTRUCTURE.  Run in the BUILD container only (it reads
/root/reference, which does not exist on the GPU box):

    python tests/golden/make_fixtures.py

It parses the INI jobs of the reference's integration suite
(/root/reference/test/test_utils.jl:77-121: network pairwise 1-3, network
advanced 1-3, raster pairwise 1-17, raster advanced 1-6, one-to-all 1-13, all-to-one 1-12), every input file
each job names, and every expected-output file `test/output_verify/<case>_*`,
and stores them as arrays in `tests/golden/reference_cases.npz`.

No reference *source* is copied -- only its test data, re-encoded.  The
readers below apply exactly the value conventions of the reference readers
so the oracle receives what the Julia code would see:

* rasters: NODATA_value -> -9999  (src/io.jl:544-549)
* .asc/.tif/.asc.gz all go through the same grid path (src/io.jl:111-120)
* text lists are stored raw (the row/col mapping needs the habitat header and
  is part of the oracle: src/io.jl:205-214)

Key layout inside the npz:  "<case>|cfg" (json), "<case>|in|<ini key>" (array),
"<case>|in|<ini key>|meta" (ncols,nrows,xll,yll,cellsize), "<case>|out|<suffix>".
"""
import gzip
import io
import json
import os
import sys

import numpy as np

REF = "/root/reference/test"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_cases.npz")

FILE_KEYS = ["habitat_file", "polygon_file", "point_file", "mask_file",
             "included_pairs_file", "variable_source_file", "source_file",
             "ground_file"]
USE_FLAG = {"polygon_file": "use_polygons", "mask_file": "use_mask",
            "included_pairs_file": "use_included_pairs",
            "variable_source_file": "use_variable_source_strengths"}


def parse_ini(path):
    """config.jl:228-242 -- skip section lines, split on the first '='."""
    cfg = {}
    with open(path) as f:
        for line in f:
            line = line.strip()
            if not line or line.startswith("[") or "=" not in line:
                continue
            k, v = line.split("=", 1)
            cfg[k.strip()] = v.strip()
    return cfg


def _open_text(path):
    if path.lower().endswith(".gz"):
        return io.TextIOWrapper(gzip.open(path, "rb"), encoding="latin-1")
    return open(path, encoding="latin-1")


def _read_bytes(path):
    if path.lower().endswith(".gz"):
        return gzip.open(path, "rb").read()
    return open(path, "rb").read()


def read_aagrid(path):
    with _open_text(path) as f:
        hdr = {}
        for _ in range(6):
            parts = f.readline().split()
            hdr[parts[0].lower()] = float(parts[1])
        data = np.loadtxt(f, dtype=np.float64, ndmin=2)
    nod = hdr.get("nodata_value", -9999.0)
    data = data.copy()
    data[data == nod] = -9999.0
    data[np.isnan(data)] = -9999.0
    meta = np.array([hdr["ncols"], hdr["nrows"], hdr["xllcorner"], hdr["yllcorner"], hdr["cellsize"]])
    assert data.shape == (int(hdr["nrows"]), int(hdr["ncols"])), (path, data.shape)
    return data, meta


def read_tif(path, like_meta=None):
    """Minimal baseline-TIFF reader (uncompressed, single band, one strip set) --
    enough for the three float64 GeoTIFFs in the reference test inputs."""
    import struct
    b = _read_bytes(path)
    bo = "<" if b[:2] == b"II" else ">"
    off = struct.unpack(bo + "I", b[4:8])[0]
    nent = struct.unpack(bo + "H", b[off:off + 2])[0]
    tags = {}
    tsize = {1: 1, 2: 1, 3: 2, 4: 4, 12: 8}
    tfmt = {1: "B", 2: "c", 3: "H", 4: "I", 12: "d"}
    for i in range(nent):
        e = b[off + 2 + 12 * i: off + 14 + 12 * i]
        tag, typ, cnt = struct.unpack(bo + "HHI", e[:8])
        nbytes = tsize[typ] * cnt
        raw = e[8:8 + nbytes] if nbytes <= 4 else b[struct.unpack(bo + "I", e[8:12])[0]:][:nbytes]
        vals = struct.unpack(bo + tfmt[typ] * cnt, raw)
        tags[tag] = vals
    w, h = tags[256][0], tags[257][0]
    assert tags[259][0] == 1 and tags[277][0] == 1, "compressed / multi-band TIFF unsupported"
    bits, fmt = tags[258][0], tags.get(339, (1,))[0]
    dt = {(64, 3): "f8", (32, 3): "f4", (32, 2): "i4", (16, 2): "i2", (8, 1): "u1",
          (16, 1): "u2", (32, 1): "u4"}[(bits, fmt)]
    chunks = [b[o:o + n] for o, n in zip(tags[273], tags[279])]
    arr = np.frombuffer(b"".join(chunks), dtype=bo + dt).astype(np.float64).reshape(h, w).copy()
    if 42113 in tags:
        nod = b"".join(tags[42113]).decode().strip("\x00 ")
        arr[arr == float(nod)] = -9999.0
    arr[np.isnan(arr)] = -9999.0
    if 33550 in tags and 33922 in tags:
        cs = float(tags[33550][0])
        xll = float(tags[33922][3])
        yll = float(tags[33922][4]) - h * cs
        meta = np.array([w, h, xll, yll, cs])
    else:
        meta = like_meta
    return arr, meta


def guess_type(path):
    """io.jl:134-158 (_guess_file_type)."""
    head = _read_bytes(path)[:4]
    if head[2:4] == b"\x2a\x00":
        return "tif"
    with _open_text(path) as f:
        hdr = f.readline()
    if hdr.lower().startswith("ncols"):
        return "aagrid"
    if hdr.startswith("min"):
        return "pairs_aagrid"
    if hdr.startswith("mode"):
        return "pairs_list"
    return "txtlist"


def load_any(path):
    t = guess_type(path)
    if t == "tif":
        a, m = read_tif(path)
        return {"kind": "grid", "data": a, "meta": m}
    if t == "aagrid":
        a, m = read_aagrid(path)
        return {"kind": "grid", "data": a, "meta": m}
    if t == "pairs_aagrid":
        with _open_text(path) as f:
            mn = float(f.readline().split()[1])
            mx = float(f.readline().split()[1])
            rows = [[float(x) for x in ln.split()] for ln in f if ln.strip()]
        return {"kind": "pairs_aagrid", "data": np.array(rows), "meta": np.array([mn, mx])}
    if t == "pairs_list":
        with _open_text(path) as f:
            mode = f.readline().split()[1]
            rows = [[float(x) for x in ln.split()] for ln in f if ln.strip()]
        return {"kind": "pairs_list_" + mode, "data": np.array(rows, dtype=np.float64).reshape(-1, 2),
                "meta": np.zeros(0)}
    with _open_text(path) as f:
        rows = [[float(x) for x in ln.split()] for ln in f if ln.strip()]
    return {"kind": "txtlist", "data": np.array(rows, dtype=np.float64), "meta": np.zeros(0)}


def load_out(path):
    if path.endswith(".asc"):
        a, _ = read_aagrid(path)
        return a
    with open(path) as f:
        rows = [[float(x) for x in ln.split()] for ln in f if ln.strip()]
    return np.array(rows, dtype=np.float64)


def cases():
    for i in range(1, 18):
        yield f"sgVerify{i}", f"input/raster/pairwise/{i}/sgVerify{i}.ini"
    for i in range(1, 7):
        yield f"mgVerify{i}", f"input/raster/advanced/{i}/mgVerify{i}.ini"
    for i in range(1, 4):
        yield f"sgNetworkVerify{i}", f"input/network/sgNetworkVerify{i}.ini"
    for i in range(1, 4):
        yield f"mgNetworkVerify{i}", f"input/network/mgNetworkVerify{i}.ini"
    for i in range(1, 14):
        yield f"oneToAllVerify{i}", f"input/raster/one_to_all/{i}/oneToAllVerify{i}.ini"
    for i in range(1, 13):
        yield f"allToOneVerify{i}", f"input/raster/all_to_one/{i}/allToOneVerify{i}.ini"


def main():
    store = {}
    verify = os.listdir(os.path.join(REF, "output_verify"))
    for name, ini in cases():
        cfg = parse_ini(os.path.join(REF, ini))
        store[f"{name}|cfg"] = np.array(json.dumps(cfg))
        for key in FILE_KEYS:
            val = cfg.get(key, "")
            flag = USE_FLAG.get(key)
            if flag and cfg.get(flag, "False") not in ("True", "true", "1"):
                continue
            p = os.path.join(REF, val)
            if not val or not os.path.isfile(p):
                continue
            d = load_any(p)
            store[f"{name}|in|{key}"] = d["data"]
            store[f"{name}|in|{key}|meta"] = d["meta"] if d["meta"] is not None else np.zeros(0)
            store[f"{name}|in|{key}|kind"] = np.array(d["kind"])
        pref = name + "_"
        for fn in sorted(verify):
            if not fn.startswith(pref) or fn.endswith(".ini"):
                continue
            suffix = fn[len(pref):]
            store[f"{name}|out|{suffix}"] = load_out(os.path.join(REF, "output_verify", fn))
    np.savez_compressed(OUT, **store)
    print(f"wrote {OUT}: {len(store)} arrays, {os.path.getsize(OUT)/1e6:.2f} MB")


if __name__ == "__main__":
    sys.exit(main())
