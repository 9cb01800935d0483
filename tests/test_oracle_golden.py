"""Pin the CPU oracle against the reference's own golden vectors for the path
(reference harness test/test_utils.jl:62-226, tolerances :72-73,
:147, :196, :217-226).  Only files the *current* reference would regenerate are
compared, exactly like `compare_all_output` (it iterates over generated files);
stale goldens of pairs that are no longer solved are ignored."""
import numpy as np
import pytest

from oracle import circuitscape_oracle as co

TOL = 1e-6


def _pair_key(name):
    parts = name.rsplit(".", 1)[0].split("_")
    try:
        return int(parts[-2]), int(parts[-1])
    except (ValueError, IndexError):
        return None


def check_resistances(x, r):
    assert x.shape == r.shape
    assert np.all(np.abs(x - r) <= np.sqrt(TOL))          # test_utils.jl:147
    # our own, much tighter, bar for the direct oracle
    # (goldens are printed with 10 digits and came from rtol-1e-6 iterative solves)
    assert np.abs(x[1:, 1:] - r[1:, 1:]).max() <= 1e-7 * max(1.0, x[1:, 1:].max())


@pytest.mark.parametrize("i", range(1, 18))
@pytest.mark.parametrize("solver", ["direct", "cholmod"])
def test_raster_pairwise(golden, i, solver):
    cfg, inp, exp = co.load_case(golden, f"sgVerify{i}")
    r = co.raster_pairwise(cfg, inp, solver)
    check_resistances(exp["resistances.out"], r.resistances)
    n = 0
    for (a, b), m in r.curmaps.items():
        assert np.sum((m - exp[f"curmap_{a}_{b}.asc"]) ** 2) < TOL   # test_utils.jl:196
        n += 1
    for (a, b), m in r.voltmaps.items():
        assert np.sum((m - exp[f"voltmap_{a}_{b}.asc"]) ** 2) < TOL
        n += 1
    if "cum_curmap.asc" in exp:
        assert np.sum((r.cum_curmap - exp["cum_curmap.asc"]) ** 2) < TOL
    if "max_curmap.asc" in exp:
        assert np.sum((r.max_curmap - exp["max_curmap.asc"]) ** 2) < TOL
    # every golden map that is not all-zero / stale must have been produced
    want = [k for k in exp if k.startswith(("curmap_", "voltmap_")) and _pair_key(k)]
    x = exp["resistances.out"]
    ids = [int(t) for t in x[0, 1:]]

    def golden_R(key):
        a, b = key
        return x[1 + ids.index(a), 1 + ids.index(b)] if a in ids and b in ids else -1

    # (pairs whose golden resistance is 0 / -1 are never post-processed by the
    #  current reference: core.jl:209-211,221 -- their golden maps are stale)
    missing = [k for k in want
               if _pair_key(k) not in (r.curmaps if k.startswith("cur") else r.voltmaps)
               and golden_R(_pair_key(k)) > 0]
    assert not missing


def _sorted_rows(a):
    return a[np.lexsort(a.T[::-1])]


@pytest.mark.parametrize("i", range(1, 4))
def test_network_pairwise(golden, i):
    cfg, inp, exp = co.load_case(golden, f"sgNetworkVerify{i}")
    r = co.network_pairwise(cfg, inp, "direct")
    x = exp["resistances.out"]
    assert np.all(x[1:, 0] + 1 == r.resistances[1:, 0])    # test_utils.jl:84-86
    assert np.all(np.abs(x[1:, 1:] - r.resistances[1:, 1:]) <= np.sqrt(TOL))
    checked = 0
    for (a, b), (nodes, cur) in r.curmaps.items():
        v = exp[f"node_currents_{a - 1}_{b - 1}.txt"].copy()
        v[:, 0] += 1
        mine = np.column_stack([nodes, cur])
        assert np.sum((_sorted_rows(mine) - _sorted_rows(v)) ** 2) < TOL
        gr, gc, val = r.branch[(a, b)]
        keep = ~np.isclose(val, 0.0, atol=1e-6)              # out.jl:117-124
        mine = np.column_stack([gr, gc, val])[keep]
        v = exp[f"branch_currents_{a - 1}_{b - 1}.txt"].copy()
        v[:, :2] += 1
        assert mine.shape == v.shape
        assert np.sum((_sorted_rows(mine) - _sorted_rows(v)) ** 2) < TOL
        key = f"voltages_{a - 1}_{b - 1}.txt"
        if key in exp:
            v = exp[key].copy()
            v[:, 0] += 1
            nodes_v, vv = r.voltmaps[(a, b)]
            assert np.sum((np.column_stack([nodes_v, vv]) - v) ** 2) < TOL
        checked += 1
    assert checked > 0
    v = exp["node_currents_cum.txt"]
    assert np.sum((r.cum_node - v[:, 1]) ** 2) < TOL


@pytest.mark.parametrize("i", range(1, 7))
def test_raster_advanced(golden, i):
    cfg, inp, exp = co.load_case(golden, f"mgVerify{i}")
    r = co.raster_advanced(cfg, inp, "direct")
    if "curmap.asc" in exp:
        assert np.sum((r.curmap - exp["curmap.asc"]) ** 2) < TOL
    if "voltmap.asc" in exp:
        assert np.sum((r.voltmap - exp["voltmap.asc"]) ** 2) < TOL


@pytest.mark.parametrize("i", range(1, 4))
def test_network_advanced(golden, i):
    cfg, inp, exp = co.load_case(golden, f"mgNetworkVerify{i}")
    r = co.network_advanced(cfg, inp, "direct")
    x = exp["voltages.txt"].copy()
    x[:, 0] += 1
    mine = np.column_stack([np.arange(1, len(r.voltages) + 1), r.voltages])
    assert np.all(np.abs(x - mine) <= np.sqrt(TOL))           # test_utils.jl:93-96
    v = exp["node_currents.txt"].copy()
    v[:, 0] += 1
    mine = np.column_stack([np.arange(1, len(r.voltages) + 1), r.node_currents])
    assert np.sum((_sorted_rows(mine) - _sorted_rows(v)) ** 2) < TOL
    gr, gc, val = r.branch
    keep = ~np.isclose(val, 0.0, atol=1e-6)
    mine = np.column_stack([gr + 1, gc + 1, val])[keep]
    v = exp["branch_currents.txt"].copy()
    v[:, :2] += 1
    assert mine.shape == v.shape
    assert np.sum((_sorted_rows(mine) - _sorted_rows(v)) ** 2) < TOL


@pytest.mark.parametrize("i", [1, 2, 4, 9, 12, 14, 16, 17])
def test_cg_amg_raster_pairwise(golden, i):
    """the oracle's reference-behaviour mode (SA-AMG-preconditioned CG, rtol 1e-6,
    src/core.jl:161-167,639) reproduces the goldens within the reference's bars."""
    cfg, inp, exp = co.load_case(golden, f"sgVerify{i}")
    r = co.raster_pairwise(cfg, inp, "cg+amg")
    x = exp["resistances.out"]
    assert np.all(np.abs(x - r.resistances) <= np.sqrt(TOL))
    for (a, b), m in r.curmaps.items():
        assert np.sum((m - exp[f"curmap_{a}_{b}.asc"]) ** 2) < TOL
    if "cum_curmap.asc" in exp:
        assert np.sum((r.cum_curmap - exp["cum_curmap.asc"]) ** 2) < TOL


def test_cg_amg_advanced(golden):
    cfg, inp, exp = co.load_case(golden, "mgVerify1")
    r = co.raster_advanced(cfg, inp, "cg+amg")
    assert np.sum((r.curmap - exp["curmap.asc"]) ** 2) < TOL
    assert np.sum((r.voltmap - exp["voltmap.asc"]) ** 2) < TOL


@pytest.mark.parametrize("name", [f"oneToAllVerify{i}" for i in range(1, 14)] +
                         [f"allToOneVerify{i}" for i in range(1, 13)])
def test_one_to_all_and_all_to_one(golden, name):
    """raster/onetoall.jl restated (next row of the scope table): resistances of every case
    and every map the current reference writes (test/test_utils.jl:123-139)."""
    cfg, inp, exp = co.load_case(golden, name)
    r = co.raster_one_to_all(cfg, inp, "direct")
    x = exp["resistances.out"]
    assert x.shape == r.resistances.shape
    assert np.all(np.abs(x - r.resistances) <= np.sqrt(TOL))
    fl = co.cfg_flags(cfg)
    for n, m in r.curmaps.items():
        assert np.sum((m - exp[f"curmap_{n}.asc"]) ** 2) < TOL
    for n, m in r.voltmaps.items():
        assert np.sum((m - exp[f"voltmap_{n}.asc"]) ** 2) < TOL
    if (fl["write_cur_maps"] or fl["write_cum_cur_map_only"]) and "cum_curmap.asc" in exp:
        assert np.sum((r.cum_curmap - exp["cum_curmap.asc"]) ** 2) < TOL
    if fl["write_max_cur_maps"] and "max_curmap.asc" in exp:
        assert np.sum((r.max_curmap - exp["max_curmap.asc"]) ** 2) < TOL
