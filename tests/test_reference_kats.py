"""Known-answer tests the reference keeps OUTSIDE its golden-file suite, restated here as in-memory
cases (no file IO) and run against the oracle, against the product's host driver on a CPU test
double, and (gpu-marked) through the CUDA library:

  /root/reference/test/issue341.jl:7-442     7 include / exclude scenarios (2 and 3 are sgVerify17 /
                                             sgVerify13, already in tests/golden)
  /root/reference/test/internal.jl:130-135   resolve_conflicts incl. the `:rmvall` quirk
  /root/reference/test/internal.jl:137-175   construct_graph stencil weights
  /root/reference/test/internal.jl:179-200   exact Laplacians of the 2x2 and 3x3 model problems

The scenario rasters are typed in from the Julia sources (5x5 / 6x6 all-ones resistance rasters,
`connect_four_neighbors_only = True`, `connect_using_avg_resistances = True`)."""
import numpy as np
import pytest
import scipy.sparse as sp

import circuitscape_b200 as cb
from circuitscape_b200 import core, graph
from circuitscape_b200 import solver as S
from oracle import circuitscape_oracle as co

from . import cases
from .fake_factor import FakeFactor

CFG = {
    "data_type": "raster", "scenario": "pairwise", "habitat_map_is_resistances": "True",
    "use_included_pairs": "True", "connect_four_neighbors_only": "True",
    "connect_using_avg_resistances": "True", "write_cur_maps": "False", "write_volt_maps": "False",
    "write_cum_cur_map_only": "False", "write_max_cur_maps": "False", "log_transform_maps": "False",
    "set_null_currents_to_nodata": "False", "set_null_voltages_to_nodata": "False", "solver": "cg+amg",
    "use_polygons": "False", "use_mask": "False",
}


def _grid(a):
    a = np.asarray(a, dtype=np.float64)
    meta = np.array([a.shape[1], a.shape[0], 0.0, 0.0, 1.0])       # ncols nrows xll yll cellsize
    return a, meta


def scenario(points, mode, pairs):
    pts, meta = _grid(points)
    cell = np.ones_like(pts)
    inp = {
        "habitat_file": ("grid", cell, meta),
        "point_file": ("grid", pts, meta),
        "included_pairs_file": (f"pairs_list_{mode}", np.array(pairs, dtype=np.float64).reshape(-1, 2), np.zeros(0)),
    }
    return dict(CFG), inp


P5 = [[1, 0, 0, 0, 2], [0] * 5, [0] * 5, [0] * 5, [3, 0, 0, 0, 0]]
P5_4 = [[1, 0, 0, 0, 2], [0] * 5, [0] * 5, [0] * 5, [3, 0, 0, 0, 4]]
P6 = [[1, 1, 0, 0, 2, 2], [0] * 6, [0] * 6, [0] * 6, [0] * 6, [3, 0, 0, 0, 0, 0]]

# name -> (points, mode, pairs, expected ids, solved (i, j) in header coordinates, excluded (i, j))
ISSUE341 = {
    "1_include_points": (P5, "include", [[1, 2]], [1, 2], [(1, 2)], []),                 # issue341.jl:7-86
    "4_include_regions": (P6, "include", [[1, 2]], [1, 2], [(1, 2)], []),                # :104-185
    "5_exclude_points": (P5, "exclude", [[1, 3]], [1, 2, 3], [(1, 2), (2, 3)], [(1, 3)]),  # :187-268
    "6_exclude_two": (P5_4, "exclude", [[1, 3], [2, 4]], [1, 2, 3, 4],
                      [(1, 2), (1, 4), (2, 3), (3, 4)], [(1, 3), (2, 4)]),               # :270-355
    "7_exclude_regions": (P6, "exclude", [[1, 3]], [1, 2, 3], [(1, 2), (2, 3)], [(1, 3)]),  # :357-442
}


def check_issue341(R, ids, solved, excluded):
    assert R.shape == (len(ids) + 1, len(ids) + 1)
    assert list(R[0, 1:]) == ids and list(R[1:, 0]) == ids
    for a, b in solved:
        assert R[a, b] > 0 and R[a, b] == R[b, a], (a, b)
    for a, b in excluded:
        assert R[a, b] == -1 and R[b, a] == -1, (a, b)
    assert np.all(np.diag(R)[1:] == 0)


def product_resistances(cfg, inp, solver):
    """the product's host driver on the same in-memory inputs (what cases.raster_pairwise_problem
    does for a packed golden case)"""
    flags = cb.Flags.from_cfg(cfg)
    fl = co.cfg_flags(cfg)
    cellmap, polymap, meta, inc = co.load_raster_inputs(cfg, inp)
    pk = inp["point_file"]
    points_rc = co.read_point_map(pk[0], pk[1], meta)
    exclude = set()
    if inc is not None:
        points_rc, exclude = co.generate_exclude_pairs(points_rc, inc)
    if len(points_rc[0]) == len(np.unique(points_rc[2])):
        nodemap = graph.construct_node_map(cellmap, polymap)
        G = graph.laplacian(graph.construct_graph(cellmap, nodemap, fl["avg_res"], fl["four_neighbors"]))
        prob = cb.GraphProblem(G, graph.connected_components(G), nodemap[points_rc[0] - 1, points_rc[1] - 1],
                               points_rc[2], exclude, nodemap, polymap, cellmap, solver)
        return cb.single_ground_all_pairs(prob, flags).resistances
    pts = list(dict.fromkeys(int(p) for p in points_rc[2]))           # focal regions: one graph per pair
    n = len(pts)
    R = -np.ones((n, n))
    for i in range(n):
        for j in range(i + 1, n):
            p1, p2 = pts[i], pts[j]
            if (p1, p2) in exclude or (p2, p1) in exclude:
                continue
            newpoly = co.create_new_polymap(cellmap, polymap, points_rc, p1, p2)
            nodemap = graph.construct_node_map(cellmap, newpoly)
            G = graph.laplacian(graph.construct_graph(cellmap, nodemap, fl["avg_res"], fl["four_neighbors"]))
            x = int(np.nonzero(points_rc[2] == p1)[0][0])
            y = int(np.nonzero(points_rc[2] == p2)[0][0])
            nodes = np.array([nodemap[points_rc[0][x] - 1, points_rc[1][x] - 1],
                              nodemap[points_rc[0][y] - 1, points_rc[1][y] - 1]])
            prob = cb.GraphProblem(G, graph.connected_components(G), nodes, np.array([p1, p2]), set(), nodemap,
                                   newpoly, cellmap, solver)
            R[i, j] = R[j, i] = cb.single_ground_all_pairs(prob, flags).resistances[1, 2]
    np.fill_diagonal(R, 0.0)
    full = np.zeros((n + 1, n + 1))
    full[0, 1:] = pts
    full[1:, 0] = pts
    full[1:, 1:] = R
    return full


@pytest.mark.parametrize("name", sorted(ISSUE341))
def test_issue341_oracle(name):
    pts, mode, pairs, ids, solved, excluded = ISSUE341[name]
    cfg, inp = scenario(pts, mode, pairs)
    check_issue341(co.raster_pairwise(cfg, inp).resistances, ids, solved, excluded)


@pytest.mark.parametrize("name", sorted(ISSUE341))
def test_issue341_product_host_driver(name, monkeypatch):
    monkeypatch.setattr(S, "construct_cholesky_factor", lambda m, s, **kw: FakeFactor(m, s, **kw))
    pts, mode, pairs, ids, solved, excluded = ISSUE341[name]
    cfg, inp = scenario(pts, mode, pairs)
    R = product_resistances(cfg, inp, cb.CUDASolver())
    check_issue341(R, ids, solved, excluded)
    Ro = co.raster_pairwise(cfg, inp).resistances
    assert np.abs(R - Ro).max() <= 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(ISSUE341))
def test_issue341_cuda(name):
    pts, mode, pairs, ids, solved, excluded = ISSUE341[name]
    cfg, inp = scenario(pts, mode, pairs)
    R = product_resistances(cfg, inp, cb.CUDASolver(rtol=1e-8))
    check_issue341(R, ids, solved, excluded)
    Ro = co.raster_pairwise(cfg, inp).resistances
    assert np.abs(R - Ro).max() <= 1e-6 * Ro.max()


def test_issue341_scenario1_values():
    """5x5 unit grid, 4 neighbours: R(corner, opposite corner along an edge) is a property of the grid;
    the included pair is solved and the pruned focal point leaves a 3x3 result (issue341.jl:79-85)."""
    cfg, inp = scenario(P5, "include", [[1, 2]])
    R = co.raster_pairwise(cfg, inp).resistances
    assert R.shape == (3, 3) and R[0, 1] == 1.0 and R[0, 2] == 2.0 and R[1, 2] > 0


# ---- test/internal.jl:130-135 ------------------------------------------------------------------
@pytest.mark.parametrize("impl", [co.resolve_conflicts, core.resolve_conflicts])
def test_resolve_conflicts_known_answers(impl):
    s, g = np.array([1.0, 0.0, 0.0]), np.array([1.0, 0.0, 0.0])
    want = {
        "rmvgnd": ([1, 0, 0], [0, 0, 0], [1, 0, 0]),
        "rmvsrc": ([0, 0, 0], [1, 0, 0], [1, 0, 0]),
        "keepall": ([1, 0, 0], [1, 0, 0], [1, 0, 0]),
        "rmvall": ([0, 0, 0], [1, 0, 0], [1, 0, 0]),      # the quirk: only the sources are zeroed
    }
    for policy, (ws, wg, wf) in want.items():
        out = impl(s.copy(), g.copy(), policy)
        assert list(out[0]) == ws and list(out[1]) == wg and list(out[2]) == wf, policy


# ---- test/internal.jl:137-175 ------------------------------------------------------------------
@pytest.mark.parametrize("impl", [co.construct_graph, graph.construct_graph])
def test_construct_graph_known_answers(impl):
    gmap = np.array([[0, 1, 2], [2, 0, 0], [2, 0, 2]], dtype=np.float64)
    nodemap = np.array([[0, 3, 4], [1, 0, 0], [2, 0, 5]])
    z = np.zeros((5, 5))

    def expect(pairs):
        m = z.copy()
        for (a, b), v in pairs.items():
            m[a, b] = m[b, a] = v
        return m

    cases_ = [
        ((False, True), {(0, 1): 2.0, (2, 3): 1.5}),
        ((True, True), {(0, 1): 2.0, (2, 3): 4.0 / 3.0}),
        ((False, False), {(0, 1): 2.0, (0, 2): 1.06066, (2, 3): 1.5}),
        ((True, False), {(0, 1): 2.0, (0, 2): 0.942809, (2, 3): 4.0 / 3.0}),
    ]
    for (avg_res, four), pairs in cases_:
        A = sp.csr_matrix(impl(gmap, nodemap, avg_res, four)).toarray()
        assert np.sum((A - expect(pairs)) ** 2) < 1e-6, (avg_res, four)


# ---- test/internal.jl:179-200 ------------------------------------------------------------------
SIZE_2 = np.array([[2, -1, -1, 0], [-1, 2, 0, -1], [-1, 0, 2, -1], [0, -1, -1, 2]], dtype=np.float64)
SIZE_3 = np.array([
    [2, -1, 0, -1, 0, 0, 0, 0, 0], [-1, 3, -1, 0, -1, 0, 0, 0, 0], [0, -1, 2, 0, 0, -1, 0, 0, 0],
    [-1, 0, 0, 3, -1, 0, -1, 0, 0], [0, -1, 0, -1, 4, -1, 0, -1, 0], [0, 0, -1, 0, -1, 3, 0, 0, -1],
    [0, 0, 0, -1, 0, 0, 2, -1, 0], [0, 0, 0, 0, -1, 0, -1, 3, -1], [0, 0, 0, 0, 0, -1, 0, -1, 2]], dtype=np.float64)


@pytest.mark.parametrize("n,want", [(2, SIZE_2), (3, SIZE_3)])
def test_model_problem_laplacians(n, want):
    """`model_problem(n)` (src/utils.jl): the n x n unit grid, 4 neighbours, through the same
    node map / graph / laplacian! chain -- oracle, product host assembly and the full-raster stencil."""
    g = np.ones((n, n))
    for mod in (co, graph):
        nm = mod.construct_node_map(g, None)
        L = sp.csr_matrix(mod.laplacian(mod.construct_graph(g, nm, False, True))).toarray()
        assert np.array_equal(L, want), mod.__name__
    L2 = graph.stencil_laplacian_from_conductance(g, four_neighbors=True).toarray()
    assert np.array_equal(L2, want)


@pytest.mark.gpu
@pytest.mark.parametrize("n,want", [(2, SIZE_2), (3, SIZE_3)])
def test_model_problem_laplacians_device_assembly(n, want):
    with cb.B200Factor.from_raster(np.ones((n, n)), cb.CUDASolver(precond="jacobi"), four_neighbors=True) as f:
        assert np.array_equal(f.get_csr().toarray(), want)
