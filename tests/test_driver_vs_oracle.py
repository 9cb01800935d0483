"""Randomised host-driver parity: the product's pairwise driver (circuitscape_b200/core.py, with the
CPU test double standing in for the device) against the oracle's independent driver on random small
rasters -- holes, short-circuit polygons, several focal ids on one node, excluded pairs, shortcut and
map-writing modes, superposition on/off.  Complements the fixed golden cases."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

import circuitscape_b200 as cb
from circuitscape_b200 import graph
from circuitscape_b200 import solver as S
from oracle import circuitscape_oracle as co

from .fake_factor import FakeFactor


@pytest.fixture(autouse=True)
def fake_device(monkeypatch):
    monkeypatch.setattr(S, "construct_cholesky_factor", lambda m, s, **kw: FakeFactor(m, s, **kw))
    monkeypatch.setattr(S, "multiple_solve", lambda s, m, b: FakeFactor(m, s).solve_rhs(np.asarray(b))[0])


@st.composite
def problems(draw):
    nr, nc = draw(st.integers(3, 7)), draw(st.integers(3, 7))
    rng = np.random.default_rng(draw(st.integers(0, 2**31 - 1)))
    g = rng.uniform(0.2, 4.0, (nr, nc))
    g[rng.random((nr, nc)) < draw(st.sampled_from([0.0, 0.15, 0.35]))] = 0.0
    poly = None
    if draw(st.booleans()):
        poly = np.zeros((nr, nc), dtype=np.int64)
        poly[rng.random((nr, nc)) < 0.2] = 1
        poly[rng.random((nr, nc)) < 0.1] = 2
    npts = draw(st.integers(2, 5))
    cells = rng.choice(nr * nc, size=npts, replace=False)
    rows, cols = cells // nc + 1, cells % nc + 1
    ids = np.arange(1, npts + 1)
    exclude = set()
    if draw(st.booleans()) and npts >= 3:
        exclude = {(int(ids[0]), int(ids[1]))}
    maps = draw(st.sampled_from(["none", "cur", "volt+cur+max", "cum_only"]))
    return g, poly, (rows, cols, ids), exclude, maps, draw(st.booleans()), draw(st.booleans())


@settings(max_examples=120, deadline=None, derandomize=True, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(problems(), st.booleans())
def test_pairwise_driver_matches_oracle(p, superpose):
    g, poly, (rows, cols, ids), exclude, maps, four, avg_res = p
    cfg = {"write_cur_maps": str(maps in ("cur", "volt+cur+max")), "write_volt_maps": str(maps == "volt+cur+max"),
           "write_max_cur_maps": str(maps == "volt+cur+max"), "write_cum_cur_map_only": str(maps == "cum_only"),
           "data_type": "raster", "scenario": "pairwise"}
    fl = co.cfg_flags(cfg)
    fl["four_neighbors"], fl["avg_res"] = four, avg_res
    nodemap = graph.construct_node_map(g, poly)
    if nodemap.max() == 0:
        return
    G = graph.laplacian(graph.construct_graph(g, nodemap, avg_res, four))
    cc = graph.connected_components(G)
    points = nodemap[rows - 1, cols - 1]
    want = co.single_ground_all_pairs(co.GraphProblem(G, cc, points, ids, set(exclude), nodemap, poly, g, True), fl)
    flags = cb.Flags.from_cfg(cfg)
    got = cb.single_ground_all_pairs(cb.GraphProblem(G, cc, points, ids, set(exclude), nodemap, poly, g,
                                                     cb.CUDASolver(superpose=superpose)), flags)
    assert got.resistances.shape == want.resistances.shape
    assert np.abs(got.resistances - want.resistances).max() < 1e-8 * max(1.0, np.abs(want.resistances).max())
    assert set(got.curmaps) == set(want.curmaps) and set(got.voltmaps) == set(want.voltmaps)
    for k in want.curmaps:
        assert np.abs(got.curmaps[k] - want.curmaps[k]).max() < 1e-8
    for k in want.voltmaps:
        assert np.abs(got.voltmaps[k] - want.voltmaps[k]).max() < 1e-8
    if maps != "none":
        assert np.abs(got.cum_curmap - want.cum_curmap).max() < 1e-8
    if want.max_curmap is not None:
        assert np.abs(got.max_curmap - want.max_curmap).max() < 1e-8


@st.composite
def advanced_problems(draw):
    nr, nc = draw(st.integers(3, 7)), draw(st.integers(3, 7))
    rng = np.random.default_rng(draw(st.integers(0, 2**31 - 1)))
    g = rng.uniform(0.2, 4.0, (nr, nc))
    g[rng.random((nr, nc)) < draw(st.sampled_from([0.0, 0.2]))] = 0.0
    src = np.where(rng.random((nr, nc)) < 0.15, rng.uniform(0.5, 2.0, (nr, nc)), 0.0)
    kind = draw(st.sampled_from(["finite", "inf", "mixed"]))
    gm = np.where(rng.random((nr, nc)) < 0.15, rng.uniform(0.5, 2.0, (nr, nc)), 0.0)
    if kind == "inf":
        gm = np.where(gm != 0, np.inf, 0.0)
    elif kind == "mixed":
        gm = np.where((gm != 0) & (rng.random((nr, nc)) < 0.5), np.inf, gm)
    policy = draw(st.sampled_from(["keepall", "rmvsrc", "rmvgnd", "rmvall"]))
    return g, src, gm, policy, draw(st.booleans())


@settings(max_examples=120, deadline=None, derandomize=True, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(p=advanced_problems())
def test_advanced_driver_matches_oracle(p):
    g, src, gm, policy, four = p
    nodemap = graph.construct_node_map(g, None)
    if nodemap.max() == 0:
        return
    G = graph.laplacian(graph.construct_graph(g, nodemap, False, four))
    cc = graph.connected_components(G)
    n = G.shape[0]
    s_o, g_o, f_o = co._sources_grounds_raster(src, gm, nodemap, n, policy)
    s_p, g_p, f_p = cb.core.sources_and_grounds_from_maps(src, gm, nodemap, n, policy)
    assert np.array_equal(s_o, s_p) and np.array_equal(g_o, g_p) and np.array_equal(f_o, f_p)
    try:
        want = co.advanced_kernel(G, cc, s_o, g_o, f_o, nodemap, None, g)
    except Exception:
        return                                   # singular set-ups (source component without a path to ground)
    got = cb.advanced_kernel(cb.AdvancedProblem(G, cc, s_p, g_p, f_p, nodemap, None, g, cb.CUDASolver()),
                             cb.Flags(is_raster=True, is_advanced=True))
    assert np.abs(got.voltages - want.voltages).max() < 1e-8 * max(1.0, np.abs(want.voltages).max())
    assert np.abs(got.voltmap - want.voltmap).max() < 1e-8 * max(1.0, np.abs(want.voltmap).max())
    assert np.abs(got.curmap - want.curmap).max() < 1e-8 * max(1.0, np.abs(want.curmap).max())


@st.composite
def onetoall_problems(draw):
    nr, nc = draw(st.integers(3, 7)), draw(st.integers(3, 7))
    rng = np.random.default_rng(draw(st.integers(0, 2**31 - 1)))
    g = rng.uniform(0.2, 4.0, (nr, nc))
    g[rng.random((nr, nc)) < draw(st.sampled_from([0.0, 0.15]))] = -9999.0
    npts = draw(st.integers(2, 5))
    cells = rng.choice(nr * nc, size=npts, replace=False)
    pm = np.zeros((nr, nc))
    ids = np.arange(1, npts + 1)
    if draw(st.booleans()) and npts >= 3:
        ids[-1] = ids[0]                                  # two cells carry one focal id (a focal region)
    pm.ravel()[cells] = ids
    poly = None
    if draw(st.booleans()):
        poly = np.zeros((nr, nc))
        poly[rng.random((nr, nc)) < 0.2] = 1
        poly[rng.random((nr, nc)) < 0.1] = 2
    strengths = None
    if draw(st.booleans()):
        u = np.unique(ids)
        strengths = np.column_stack([u, rng.uniform(0.5, 3.0, len(u))])
    scenario = draw(st.sampled_from(["one-to-all", "all-to-one"]))
    maps = draw(st.sampled_from(["cur", "volt+cur+max", "cum_only"]))
    return g, pm, poly, strengths, scenario, maps, draw(st.booleans())


@settings(max_examples=120, deadline=None, derandomize=True, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(p=onetoall_problems(), batched=st.booleans())
def test_onetoall_driver_matches_oracle(p, batched):
    g, pm, poly, strengths, scenario, maps, four = p
    nr, nc = g.shape
    meta = np.array([nc, nr, 0.0, 0.0, 1.0])
    cfg = {"scenario": scenario, "data_type": "raster", "habitat_map_is_resistances": "False",
           "write_cur_maps": str(maps in ("cur", "volt+cur+max")), "write_volt_maps": str(maps == "volt+cur+max"),
           "write_max_cur_maps": str(maps == "volt+cur+max"), "write_cum_cur_map_only": str(maps == "cum_only"),
           "use_polygons": str(poly is not None), "use_variable_source_strengths": str(strengths is not None),
           "connect_four_neighbors_only": str(four)}
    inputs = {"habitat_file": ("grid", g, meta), "point_file": ("grid", pm, meta)}
    if poly is not None:
        inputs["polygon_file"] = ("grid", poly, meta)
    if strengths is not None:
        inputs["variable_source_file"] = ("txtlist", strengths, np.zeros(0))
    try:
        want = co.raster_one_to_all(cfg, inputs)
    except (ValueError, IndexError):
        return        # inputs the reference rejects too: < 2 valid focal nodes, strengths with focal regions
    cellmap, polymap, _, inc = co.load_raster_inputs(cfg, inputs)
    points_rc = co.read_point_map("grid", pm, meta)
    data = cb.RasterData(cellmap, polymap, points_rc, None if strengths is None else strengths.copy(), inc)
    flags = cb.Flags.from_cfg(cfg)
    got = cb.onetoall_kernel(data, flags, cfg, solver=cb.CUDASolver(batch_all_to_one=batched, batch_one_to_all=batched),
                             four_neighbors=four, avg_res=False)
    assert got.resistances.shape == want.resistances.shape
    assert np.abs(got.resistances - want.resistances).max() < 1e-8 * max(1.0, np.abs(want.resistances).max())
    assert set(got.curmaps) == set(want.curmaps) and set(got.voltmaps) == set(want.voltmaps)
    for k in want.curmaps:
        assert np.abs(got.curmaps[k] - want.curmaps[k]).max() < 1e-8 * max(1.0, np.abs(want.curmaps[k]).max())
    for k in want.voltmaps:
        assert np.abs(got.voltmaps[k] - want.voltmaps[k]).max() < 1e-8 * max(1.0, np.abs(want.voltmaps[k]).max())
    assert np.abs(got.cum_curmap - want.cum_curmap).max() < 1e-8 * max(1.0, np.abs(want.cum_curmap).max())
    if want.max_curmap is not None:
        assert np.abs(got.max_curmap - want.max_curmap).max() < 1e-8 * max(1.0, np.abs(want.max_curmap).max())


@st.composite
def networks(draw):
    n = draw(st.integers(4, 12))
    rng = np.random.default_rng(draw(st.integers(0, 2**31 - 1)))
    edges = {(i, i + 1) for i in range(1, n) if rng.random() < 0.85}       # mostly a path: a few components
    for _ in range(draw(st.integers(0, 2 * n))):
        a, b = rng.integers(1, n + 1, 2)
        if a != b:
            edges.add((min(a, b), max(a, b)))
    if not edges:
        edges = {(1, 2)}
    e = np.array(sorted(edges), dtype=np.float64)
    if 1 not in e[:, :2]:
        e = np.vstack([e, [1, 2]])
    raw = np.column_stack([e, rng.uniform(0.2, 3.0, len(e))])
    nn = int(raw[:, :2].max())
    k = draw(st.integers(2, min(5, nn)))
    fp = np.sort(rng.choice(np.arange(1, nn + 1), size=k, replace=False))
    return raw, fp


@settings(max_examples=120, deadline=None, derandomize=True, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(p=networks(), superpose=st.booleans())
def test_network_pairwise_driver_matches_oracle(p, superpose):
    raw, fp = p
    cfg = {"data_type": "network", "scenario": "pairwise", "habitat_map_is_resistances": "False",
           "write_cur_maps": "True", "write_volt_maps": "True"}
    inputs = {"habitat_file": ("txtlist", raw, np.zeros(0)), "point_file": ("txtlist", fp.reshape(-1, 1), np.zeros(0))}
    want = co.network_pairwise(cfg, inputs)
    i, j, v, _ = co.load_graph(raw, False)
    G, cc = co.network_graph(i, j, v)
    flags = cb.Flags.from_cfg(cfg)
    got = cb.single_ground_all_pairs(cb.GraphProblem(G, cc, fp, fp, set(), None, None, None,
                                                     cb.CUDASolver(superpose=superpose), (i, j)), flags)
    assert np.abs(got.resistances - want.resistances).max() < 1e-8 * max(1.0, np.abs(want.resistances).max())
    assert set(got.curmaps) == set(want.curmaps)
    for k in want.curmaps:
        gn, gc_ = got.curmaps[k]
        wn, wc = want.curmaps[k]
        assert np.array_equal(np.asarray(gn), np.asarray(wn))
        assert np.abs(gc_ - wc).max() < 1e-8 * max(1.0, np.abs(wc).max())
        srt = lambda t: (lambda m: m[np.lexsort(m.T[::-1])])(np.column_stack([np.asarray(x, dtype=float) for x in t]))
        gb, wb = srt(got.branch[k]), srt(want.branch[k])      # row order is not pinned (test_utils.jl sorts too)
        assert gb.shape == wb.shape and np.abs(gb - wb).max(initial=0.0) < 1e-8
    assert np.abs(got.cum_node - want.cum_node).max() < 1e-8 * max(1.0, np.abs(want.cum_node).max())
    assert np.abs(got.cum_branch - want.cum_branch).max() < 1e-8 * max(1.0, np.abs(want.cum_branch).max())


def test_onetoall_batched_reads_value_through_the_local_node_map():
    """regression (found by the randomised comparison): a focal region whose first cell is NODATA --
    the reported value is read off the voltage raster at the source cell, where the reference's
    local node numbering differs from the matrix's; the batched path must reproduce that."""
    N = -9999.0
    g = np.array([[3.02968197, N, 3.79571608], [N, 1.90727603, N], [2.13672554, 1.50118691, N],
                  [N, 1.14983282, 3.50826361], [N, 0.25732439, 1.38680231], [1.92996258, 2.58616864, 0.621780943]])
    pm = np.array([[0, 0, 0], [0, 0, 1], [0, 0, 0], [0, 0, 0], [2, 0, 1], [3, 0, 0.]])
    meta = np.array([3, 6, 0.0, 0.0, 1.0])
    cfg = {"scenario": "one-to-all", "data_type": "raster", "habitat_map_is_resistances": "False",
           "write_cur_maps": "True", "use_polygons": "False", "connect_four_neighbors_only": "True"}
    inputs = {"habitat_file": ("grid", g, meta), "point_file": ("grid", pm, meta)}
    want = co.raster_one_to_all(cfg, inputs)
    cellmap, polymap, _, inc = co.load_raster_inputs(cfg, inputs)
    data = cb.RasterData(cellmap, polymap, co.read_point_map("grid", pm, meta), None, inc)
    for batched in (False, True):
        got = cb.onetoall_kernel(data, cb.Flags.from_cfg(cfg), cfg, solver=cb.CUDASolver(batch_one_to_all=batched),
                                 four_neighbors=True, avg_res=False)
        assert np.abs(got.resistances - want.resistances).max() < 1e-9


@settings(max_examples=120, deadline=None, derandomize=True, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(p=networks(), seed=st.integers(0, 2**31 - 1), policy=st.sampled_from(["keepall", "rmvsrc", "rmvgnd", "rmvall"]),
       kind=st.sampled_from(["finite", "inf", "mixed"]))
def test_network_advanced_driver_matches_oracle(p, seed, policy, kind):
    raw, _ = p
    rng = np.random.default_rng(seed)
    i, j, v, _ = co.load_graph(raw, False)
    G, cc = co.network_graph(i, j, v)
    n = G.shape[0]
    sources = np.where(rng.random(n) < 0.3, rng.uniform(0.5, 2.0, n), 0.0)
    grounds = np.where(rng.random(n) < 0.3, rng.uniform(0.5, 2.0, n), 0.0)
    if kind == "inf":
        grounds = np.where(grounds != 0, np.inf, 0.0)
    elif kind == "mixed":
        grounds = np.where((grounds != 0) & (rng.random(n) < 0.5), np.inf, grounds)
    s_o, g_o, f_o = co.resolve_conflicts(sources, grounds, policy)
    s_p, g_p, f_p = cb.resolve_conflicts(sources, grounds, policy)
    assert np.array_equal(s_o, s_p) and np.array_equal(g_o, g_p) and np.array_equal(f_o, f_p)
    try:
        want = co.advanced_kernel(G, cc, s_o, g_o, f_o)
    except Exception:
        return
    got = cb.advanced_kernel(cb.AdvancedProblem(G, cc, s_p, g_p, f_p, None, None, None, cb.CUDASolver()),
                             cb.Flags(is_raster=False, is_advanced=True))
    scale = max(1.0, np.abs(want.voltages).max())
    assert np.abs(got.voltages - want.voltages).max() < 1e-8 * scale
    assert np.abs(got.node_currents - want.node_currents).max() < 1e-8 * max(1.0, np.abs(want.node_currents).max())
    srt = lambda t: (lambda m: m[np.lexsort(m.T[::-1])])(np.column_stack([np.asarray(x, dtype=float) for x in t]))
    wr, wc, wv = want.branch                               # the oracle keeps 0-based rows here
    gb, wb = srt(got.branch), srt((np.asarray(wr) + 1, np.asarray(wc) + 1, wv))
    assert gb.shape == wb.shape and np.abs(gb - wb).max(initial=0.0) < 1e-8 * max(1.0, np.abs(wb).max(initial=0.0))
