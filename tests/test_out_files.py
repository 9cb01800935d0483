"""Output stage (circuitscape_b200/out.py): files written from the host driver's results are
read back and compared with the reference's golden files under the reference's own tolerances
(test/test_utils.jl:144-163, 196, 217-226).  CPU only (FakeFactor stands in for the device)."""
import os

import numpy as np
import pytest

import circuitscape_b200 as cb
from circuitscape_b200 import out as O
from circuitscape_b200 import solver as S
from oracle import circuitscape_oracle as co

from . import cases
from .fake_factor import FakeFactor

TOL = 1e-6


@pytest.fixture(autouse=True)
def fake_device(monkeypatch, request):
    if "gpu" in request.keywords:            # the gpu-marked tests write files from REAL device results
        return
    monkeypatch.setattr(S, "construct_cholesky_factor", lambda m, s, **kw: FakeFactor(m, s, **kw))
    monkeypatch.setattr(S, "multiple_solve", lambda s, m, b: FakeFactor(m, s).solve_rhs(np.asarray(b))[0])


def read_asc(path):
    with open(path) as f:
        hdr = {}
        for _ in range(6):
            k, v = f.readline().split()
            hdr[k.lower()] = float(v)
        a = np.array([[float(x) for x in ln.split()] for ln in f if ln.strip()])
    assert a.shape == (int(hdr["nrows"]), int(hdr["ncols"]))
    return a, hdr


def read_table(path):
    with open(path) as f:
        return np.array([[float(x) for x in ln.split()] for ln in f if ln.strip()])


@pytest.mark.parametrize("name", ["sgVerify1", "sgVerify3", "sgVerify12"])
def test_raster_pairwise_files(golden, name, tmp_path):
    _raster_files(golden, tmp_path, name, cb.CUDASolver())


def _raster_files(golden, tmp_path, name, solver):
    r, exp = cases.run_raster_pairwise(golden, name, solver)
    cfg, inp, _ = co.load_case(golden, name)
    shape = inp["habitat_file"][1].shape
    meta = O.RasterMeta(ncols=shape[1], nrows=shape[0], xllcorner=3.5, yllcorner=-2.0, cellsize=0.25)
    of = str(tmp_path / "out" / f"{name}.out")
    flags = cb.Flags.from_cfg(cfg)
    written = O.write_pairwise_outputs(r, of, meta, write_cum=True, write_max=flags.outputflags.write_max_cur_maps)
    assert all(os.path.exists(p) for p in written)
    pref = of[:-4]
    res = read_table(pref + "_resistances.out")
    assert np.all(np.abs(res - exp["resistances.out"]) <= np.sqrt(TOL))
    c3 = read_table(pref + "_resistances_3columns.out")
    if "resistances_3columns.out" in exp and exp["resistances_3columns.out"].shape == c3.shape:
        assert np.all(np.abs(c3 - exp["resistances_3columns.out"]) <= np.sqrt(TOL))
    n = 0
    for key, gold in exp.items():
        if not key.endswith(".asc"):
            continue
        path = f"{pref}_{key}"
        if not os.path.exists(path):
            continue                                     # stale goldens the current reference does not write
        a, hdr = read_asc(path)
        assert hdr["nodata_value"] == -9999 and hdr["cellsize"] == 0.25 and hdr["xllcorner"] == 3.5
        assert np.sum((a - gold) ** 2) < TOL
        n += 1
    assert n >= 1


def test_network_pairwise_files(golden, tmp_path):
    _network_files(golden, tmp_path, cb.CUDASolver())


def _network_files(golden, tmp_path, solver):
    prob, flags, exp = cases.network_pairwise_problem(golden, "sgNetworkVerify1", solver)
    r = cb.single_ground_all_pairs(prob, flags)
    of = str(tmp_path / "net.out")
    O.write_pairwise_outputs(r, of, None)
    pref = of[:-4]
    res = read_table(pref + "_resistances.out")
    x = exp["resistances.out"]
    assert np.all(np.abs(x[1:, 1:] - res[1:, 1:]) <= np.sqrt(TOL))
    (a, b) = next(iter(r.curmaps))
    nodes = read_table(f"{pref}_node_currents_{a}_{b}.txt")
    v = exp[f"node_currents_{a - 1}_{b - 1}.txt"].copy(); v[:, 0] += 1
    assert np.sum((cases.sorted_rows(nodes) - cases.sorted_rows(v)) ** 2) < TOL
    br = read_table(f"{pref}_branch_currents_{a}_{b}.txt")
    v = exp[f"branch_currents_{a - 1}_{b - 1}.txt"].copy(); v[:, :2] += 1
    assert br.shape == v.shape                          # rows within 1e-6 of zero are dropped, as in the reference
    assert np.sum((cases.sorted_rows(br) - cases.sorted_rows(v)) ** 2) < TOL
    if r.voltmaps:
        assert os.path.exists(f"{pref}_voltages_{a}_{b}.txt")


def test_grid_names_and_number_format(tmp_path):
    assert O.grid_filename("x/y.out", "_1_2").endswith("x/y_curmap_1_2.asc")
    assert O.grid_filename("x/y.out", "", cum=True, maxmap=True).endswith("y_cum_curmap.asc")
    assert O.grid_filename("x/y.out", "", maxmap=True, voltage=True).endswith("y_max_curmap.asc")
    assert O.grid_filename("x/y.out", "_3", voltage=True).endswith("y_voltmap_3.asc")
    m = O.RasterMeta(ncols=3, nrows=2)
    a = np.array([[0.1, -9999.0, 3.0], [1e-17, 2.5e10, 7.0]])
    p = O.write_asc(str(tmp_path / "g.asc"), a, m)
    b, _ = read_asc(p)
    assert np.array_equal(a, b)                         # repr-exact floats survive the round trip
    with pytest.raises(ValueError):
        O.write_asc(str(tmp_path / "h.asc"), a.T, m)


# ---- the same file-level checks with results that came from the CUDA library -------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["sgVerify1", "sgVerify2", "sgVerify5", "sgVerify9", "sgVerify14", "sgVerify16"])
def test_raster_pairwise_files_from_device_results(golden, tmp_path, name):
    """device results -> the reference's file set -> read back -> goldens"""
    _raster_files(golden, tmp_path, name, cb.CUDASolver(rtol=1e-8))


@pytest.mark.gpu
def test_network_pairwise_files_from_device_results(golden, tmp_path):
    _network_files(golden, tmp_path, cb.CUDASolver(rtol=1e-8))


@pytest.mark.gpu
def test_streaming_sink_writes_the_same_files_from_device_results(golden, tmp_path):
    """maps handed to a writer sink as each batch finishes (nothing kept in memory) give the files
    the keep-everything path gives"""
    name = "sgVerify1"
    kept, _ = cases.run_raster_pairwise(golden, name, cb.CUDASolver(rtol=1e-8))
    shape = co.load_case(golden, name)[1]["habitat_file"][1].shape
    meta = O.RasterMeta(ncols=shape[1], nrows=shape[0], xllcorner=0.0, yllcorner=0.0, cellsize=1.0)
    a = str(tmp_path / "a" / f"{name}.out")
    b = str(tmp_path / "b" / f"{name}.out")
    os.makedirs(os.path.dirname(b))
    O.write_pairwise_outputs(kept, a, meta)

    class Sink:
        def voltmap(self, key, grid):
            O.write_grid(grid, f"_{key[0]}_{key[1]}", b, meta, voltage=True)

        def curmap(self, key, grid):
            O.write_grid(grid, f"_{key[0]}_{key[1]}", b, meta)

    streamed, _ = cases.run_raster_pairwise(golden, name, cb.CUDASolver(rtol=1e-8), sink=Sink())
    assert not streamed.curmaps and not streamed.voltmaps
    pa, pb = os.path.dirname(a), os.path.dirname(b)
    maps = sorted(f for f in os.listdir(pb) if f.endswith(".asc"))
    assert maps
    for f in maps:
        ga, _ = read_asc(os.path.join(pa, f))
        gb, _ = read_asc(os.path.join(pb, f))
        assert np.array_equal(ga, gb), f


# ---- GeoTIFF output (cfg.write_as_tif, src/out.jl:338,378,483-531) ---------------------------------
def read_tif(path):
    """Minimal little-endian classic-TIFF reader for what `write_tif` emits: returns (array, tags)."""
    import struct
    import zlib
    raw = open(path, "rb").read()
    assert raw[:4] == b"II*\x00"
    (ifd,) = struct.unpack_from("<I", raw, 4)
    (cnt,) = struct.unpack_from("<H", raw, ifd)
    size = {2: 1, 3: 2, 4: 4, 12: 8}
    code = {3: "H", 4: "I", 12: "d"}
    tags, prev = {}, 0
    for i in range(cnt):
        tag, typ, n = struct.unpack_from("<HHI", raw, ifd + 2 + 12 * i)
        assert tag > prev            # TIFF 6.0: entries sorted by tag
        prev = tag
        at = ifd + 2 + 12 * i + 8
        if n * size[typ] > 4:
            (at,) = struct.unpack_from("<I", raw, at)
            assert at % 2 == 0       # values start on a word boundary
        tags[tag] = raw[at:at + n] if typ == 2 else list(struct.unpack_from("<" + code[typ] * n, raw, at))
    assert struct.unpack_from("<I", raw, ifd + 2 + 12 * cnt)[0] == 0     # single IFD
    ncols, nrows, bps = tags[256][0], tags[257][0], tags[258][0]
    assert tags[339] == [3] and tags[277] == [1] and tags[262] == [1]
    parts = []
    for off, nb in zip(tags[273], tags[279]):
        b = raw[off:off + nb]
        parts.append(zlib.decompress(b) if tags[259] == [8] else b)
    a = np.frombuffer(b"".join(parts), dtype="<f8" if bps == 64 else "<f4").reshape(nrows, ncols)
    return a, tags


@pytest.mark.parametrize("compress", ["deflate", "none"])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_write_tif_round_trip(tmp_path, compress, dtype):
    rng = np.random.default_rng(7)
    a = rng.random((37, 53)).astype(dtype)
    a[3, 4] = -9999.0
    wkt = 'PROJCS["NAD83 / UTM zone 10N",GEOGCS["NAD83",DATUM["North_American_Datum_1983"]]]'
    m = O.RasterMeta(ncols=53, nrows=37, xllcorner=100.5, yllcorner=-20.0, cellsize=0.25, wkt=wkt)
    p = O.write_tif(str(tmp_path / "t.tif"), a, m, compress=compress, rows_per_strip=5)
    b, tags = read_tif(p)
    assert b.dtype == np.dtype(dtype) and np.array_equal(b, a)           # lossless
    assert len(tags[273]) == 8 and tags[278] == [5]
    assert tags[33550] == [0.25, 0.25, 0.0]                              # ModelPixelScale
    assert tags[33922] == [0.0, 0.0, 0.0, 100.5, -20.0 + 37 * 0.25, 0.0]  # tiepoint: upper-left corner
    assert tags[42113].rstrip(b"\0") == b"-9999"                         # GDAL_NODATA
    keys = tags[34735]
    assert keys[:4] == [1, 1, 0, 4] and keys[4:8] == [1024, 0, 1, 1] and keys[12:16] == [3072, 0, 1, 32767]
    assert keys[16:18] == [3073, 34737] and keys[19] == 0
    cit = tags[34737][keys[19]:keys[19] + keys[18]]
    assert cit == ("ESRI PE String = " + wkt + "|").encode()


def test_write_tif_rotated_transform_and_no_projection(tmp_path):
    a = np.arange(12, dtype=np.float64).reshape(3, 4)
    m = O.RasterMeta(ncols=4, nrows=3, transform=(10.0, 2.0, 0.5, 50.0, -0.25, -2.0))
    b, tags = read_tif(O.write_tif(str(tmp_path / "r.tif"), a, m))
    assert np.array_equal(b, a)
    assert 33550 not in tags and 33922 not in tags
    assert tags[34264] == [2.0, 0.5, 0.0, 10.0, -0.25, -2.0, 0.0, 50.0, 0, 0, 0, 0, 0, 0, 0, 1.0]
    assert tags[34735] == [1, 1, 0, 1, 1025, 0, 1, 1] and 34737 not in tags
    with pytest.raises(ValueError):
        O.write_tif(str(tmp_path / "bad.tif"), a, O.RasterMeta(ncols=5, nrows=3))
    with pytest.raises(ValueError):
        O.write_tif(str(tmp_path / "bad.tif"), a, m, compress="lzw")


def test_write_tif_is_readable_by_an_independent_decoder(tmp_path):
    Image = pytest.importorskip("PIL.Image")
    a = np.random.default_rng(1).random((40, 31)).astype(np.float32)
    m = O.RasterMeta(ncols=31, nrows=40, cellsize=30.0)
    for compress in ("deflate", "none"):
        p = O.write_tif(str(tmp_path / f"{compress}.tif"), a, m, compress=compress)
        with Image.open(p) as im:
            assert im.size == (31, 40) and im.mode == "F"
            assert np.array_equal(np.array(im), a)
            assert im.tag_v2[33550] == (30.0, 30.0, 0.0)


def test_pairwise_outputs_as_tif(golden, tmp_path):
    r, _ = cases.run_raster_pairwise(golden, "sgVerify1", cb.CUDASolver())
    _, inp, _ = co.load_case(golden, "sgVerify1")
    shape = inp["habitat_file"][1].shape
    meta = O.RasterMeta(ncols=shape[1], nrows=shape[0], xllcorner=3.5, yllcorner=-2.0, cellsize=0.25)
    asc = O.write_pairwise_outputs(r, str(tmp_path / "a" / "x.out"), meta)
    tif = O.write_pairwise_outputs(r, str(tmp_path / "t" / "x.out"), meta, write_as_tif=True)
    assert [os.path.basename(p).replace(".tif", ".asc") for p in tif] == [os.path.basename(p) for p in asc]
    n = 0
    for pa, pt in zip(asc, tif):
        if pa.endswith(".asc"):
            assert pt.endswith(".tif")
            assert np.array_equal(read_tif(pt)[0], read_asc(pa)[0])       # same pixels in both formats
            n += 1
    assert n >= 2
