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


@settings(max_examples=120, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
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


@settings(max_examples=120, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
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
