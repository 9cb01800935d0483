"""Host assembly of the product (circuitscape_b200/graph.py) against the oracle's independent
restatement on random small rasters -- holes, NODATA, short-circuit polygons, both averaging
rules, 4/8 neighbours (src/raster/pairwise.jl:271-367, src/core.jl:608-624)."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from circuitscape_b200 import graph
from oracle import circuitscape_oracle as co


@st.composite
def rasters(draw):
    nr = draw(st.integers(1, 9))
    nc = draw(st.integers(1, 9))
    seed = draw(st.integers(0, 2**31 - 1))
    rng = np.random.default_rng(seed)
    g = rng.uniform(0.1, 5.0, (nr, nc))
    g[rng.random((nr, nc)) < draw(st.sampled_from([0.0, 0.2, 0.5]))] = 0.0
    poly = None
    if draw(st.booleans()):
        poly = np.zeros((nr, nc), dtype=np.int64)
        for pid in range(1, draw(st.integers(1, 3)) + 1):
            poly[rng.random((nr, nc)) < 0.2] = pid
    return g, poly


@settings(max_examples=150, deadline=None, derandomize=True)
@given(rasters(), st.booleans(), st.booleans())
def test_node_map_graph_and_laplacian_agree(rp, avg_res, four):
    g, poly = rp
    nm_p = graph.construct_node_map(g, poly)
    nm_o = co.construct_node_map(g, poly)
    assert np.array_equal(nm_p, nm_o)
    if nm_p.max() == 0:
        return
    a_p = graph.construct_graph(g, nm_p, avg_res, four)
    a_o = co.construct_graph(g, nm_o, avg_res, four)
    assert a_p.shape == a_o.shape
    assert abs(a_p - a_o).max() <= 1e-14 * max(1.0, abs(a_o).max())
    L_p = graph.laplacian(a_p)
    L_o = co.laplacian(a_o)
    assert abs(L_p - L_o).max() <= 1e-13 * max(1.0, abs(L_o).max())
    cc_p = sorted(tuple(c) for c in graph.connected_components(L_p))
    cc_o = sorted(tuple(c) for c in co.connected_components(L_o))
    assert cc_p == cc_o


@pytest.mark.parametrize("shape", [(1, 1), (1, 5), (6, 1), (7, 4)])
@pytest.mark.parametrize("four", [False, True])
def test_full_raster_stencil_matches_general_path(shape, four):
    rng = np.random.default_rng(3)
    g = rng.uniform(0.5, 2.0, shape)
    L1 = graph.stencil_laplacian_from_conductance(g, four, False)
    nm = graph.construct_node_map(g, None)
    if g.size == 1:
        assert L1.shape == (1, 1)
        return
    L2 = graph.laplacian(graph.construct_graph(g, nm, False, four))
    assert abs(L1 - L2).max() < 1e-14
