"""Shared helpers: build the product's GraphProblem / AdvancedProblem from the
packed reference fixtures (input parsing = oracle front end, test infrastructure)
and compare outputs with the reference's own tolerances (test/test_utils.jl)."""
import numpy as np

import circuitscape_b200 as cb
from circuitscape_b200 import graph
from oracle import circuitscape_oracle as co

TOL = 1e-6


def raster_pairwise_problem(golden, name, solver):
    """-> (list of (GraphProblem, pt1, pt2) , flags, expected).  One problem for the
    plain path; one per focal-region pair for the polygon path
    (src/raster/pairwise.jl:72-135)."""
    cfg, inp, exp = co.load_case(golden, name)
    flags = cb.Flags.from_cfg(cfg)
    fl = co.cfg_flags(cfg)
    cellmap, polymap, meta, inc = co.load_raster_inputs(cfg, inp)
    pk = inp["point_file"]
    points_rc = co.read_point_map(pk[0], pk[1], meta)
    probs = []
    if len(points_rc[0]) != len(np.unique(points_rc[2])):
        exclude = set()
        if inc is not None:
            points_rc, exclude = co.generate_exclude_pairs(points_rc, inc)
        pts = list(dict.fromkeys(int(p) for p in points_rc[2]))
        for i in range(len(pts)):
            for j in range(i + 1, len(pts)):
                p1, p2 = pts[i], pts[j]
                if (p1, p2) in exclude or (p2, p1) in exclude:
                    continue
                newpoly = co.create_new_polymap(cellmap, polymap, points_rc, p1, p2)
                nodemap = graph.construct_node_map(cellmap, newpoly)
                G = graph.laplacian(graph.construct_graph(cellmap, nodemap, fl["avg_res"], fl["four_neighbors"]))
                cc = graph.connected_components(G)
                x = int(np.nonzero(points_rc[2] == p1)[0][0])
                y = int(np.nonzero(points_rc[2] == p2)[0][0])
                pts_nodes = np.array([nodemap[points_rc[0][x] - 1, points_rc[1][x] - 1],
                                      nodemap[points_rc[0][y] - 1, points_rc[1][y] - 1]])
                probs.append((cb.GraphProblem(G, cc, pts_nodes, np.array([p1, p2]), set(), nodemap,
                                              newpoly, cellmap, solver), p1, p2))
        return probs, flags, exp, pts
    exclude = set()
    if inc is not None:
        points_rc, exclude = co.generate_exclude_pairs(points_rc, inc)
    nodemap = graph.construct_node_map(cellmap, polymap)
    G = graph.laplacian(graph.construct_graph(cellmap, nodemap, fl["avg_res"], fl["four_neighbors"]))
    cc = graph.connected_components(G)
    points = nodemap[points_rc[0] - 1, points_rc[1] - 1]
    probs.append((cb.GraphProblem(G, cc, points, points_rc[2], exclude, nodemap, polymap, cellmap, solver),
                  None, None))
    return probs, flags, exp, None


def run_raster_pairwise(golden, name, solver, sink=None):
    """Returns an object with .resistances/.curmaps/.voltmaps/.cum_curmap/.max_curmap."""
    probs, flags, exp, pts = raster_pairwise_problem(golden, name, solver)
    if pts is None:
        return cb.single_ground_all_pairs(probs[0][0], flags, sink=sink), exp
    n = len(pts)
    R = -np.ones((n, n))
    merged = None
    for prob, p1, p2 in probs:
        r = cb.single_ground_all_pairs(prob, flags)
        i, j = pts.index(p1), pts.index(p2)
        R[i, j] = R[j, i] = r.resistances[1, 2]
        if merged is None:
            merged = r
        else:
            merged.voltmaps.update(r.voltmaps)
            merged.curmaps.update(r.curmaps)
            merged.cum_curmap = merged.cum_curmap + r.cum_curmap
            if merged.max_curmap is not None:
                merged.max_curmap = np.maximum(merged.max_curmap, r.max_curmap)
            merged.num_solves += r.num_solves
    np.fill_diagonal(R, 0.0)
    full = np.zeros((n + 1, n + 1))
    full[0, 1:] = pts
    full[1:, 0] = pts
    full[1:, 1:] = R
    merged.resistances = full
    return merged, exp


def check_raster_pairwise(r, exp, rel=1e-6, map_tol=TOL):
    x = exp["resistances.out"]
    assert x.shape == r.resistances.shape
    assert np.all(np.abs(x - r.resistances) <= np.sqrt(TOL))          # reference bar
    assert np.abs(x[1:, 1:] - r.resistances[1:, 1:]).max() <= rel * max(1.0, x[1:, 1:].max())
    for (a, b), m in r.curmaps.items():
        assert np.sum((m - exp[f"curmap_{a}_{b}.asc"]) ** 2) < map_tol
    for (a, b), m in r.voltmaps.items():
        assert np.sum((m - exp[f"voltmap_{a}_{b}.asc"]) ** 2) < map_tol
    if "cum_curmap.asc" in exp:
        assert np.sum((r.cum_curmap - exp["cum_curmap.asc"]) ** 2) < map_tol
    if "max_curmap.asc" in exp:
        assert np.sum((r.max_curmap - exp["max_curmap.asc"]) ** 2) < map_tol


def network_pairwise_problem(golden, name, solver):
    cfg, inp, exp = co.load_case(golden, name)
    flags = cb.Flags.from_cfg(cfg)
    i, j, v, _ = co.load_graph(inp["habitat_file"][1], co.cfg_bool(cfg, "habitat_map_is_resistances", "True"))
    fp = np.asarray(inp["point_file"][1]).ravel().astype(np.int64)
    if fp.min() == 0:
        fp = fp + 1
    import scipy.sparse as sp
    m = int(max(i.max(), j.max()))
    A = sp.coo_matrix((v, (i - 1, j - 1)), shape=(m, m)).tocsr()
    A = (A + A.T).tocsr()
    G = graph.laplacian(A)
    cc = graph.connected_components(A)
    return cb.GraphProblem(G, cc, fp, fp, set(), None, None, None, solver, (i, j)), flags, exp


def sorted_rows(a):
    return a[np.lexsort(a.T[::-1])]


def check_network_pairwise(r, exp, map_tol=TOL):
    x = exp["resistances.out"]
    assert np.all(x[1:, 0] + 1 == r.resistances[1:, 0])
    assert np.all(np.abs(x[1:, 1:] - r.resistances[1:, 1:]) <= np.sqrt(TOL))
    assert r.curmaps
    for (a, b), (nodes, cur) in r.curmaps.items():
        v = exp[f"node_currents_{a - 1}_{b - 1}.txt"].copy()
        v[:, 0] += 1
        assert np.sum((sorted_rows(np.column_stack([nodes, cur])) - sorted_rows(v)) ** 2) < map_tol
        gr, gc, val = r.branch[(a, b)]
        keep = ~np.isclose(val, 0.0, atol=1e-6)
        mine = np.column_stack([gr, gc, val])[keep]
        v = exp[f"branch_currents_{a - 1}_{b - 1}.txt"].copy()
        v[:, :2] += 1
        assert mine.shape == v.shape
        assert np.sum((sorted_rows(mine) - sorted_rows(v)) ** 2) < map_tol
    v = exp["node_currents_cum.txt"]
    assert np.sum((r.cum_node - v[:, 1]) ** 2) < map_tol


def advanced_problem(golden, name, solver):
    cfg, inp, exp = co.load_case(golden, name)
    flags = cb.Flags.from_cfg(cfg)
    fl = co.cfg_flags(cfg)
    if flags.is_raster:
        cellmap, polymap, meta, _ = co.load_raster_inputs(cfg, inp)
        source_map, ground_map = co.read_source_and_ground_maps(cfg, inp, meta)
        nodemap = graph.construct_node_map(cellmap, polymap)
        G = graph.laplacian(graph.construct_graph(cellmap, nodemap, fl["avg_res"], fl["four_neighbors"]))
        cc = graph.connected_components(G)
        s, g, f = co._sources_grounds_raster(source_map, ground_map, nodemap, G.shape[0], fl["policy"])
        return cb.AdvancedProblem(G, cc, s, g, f, nodemap, polymap, cellmap, solver), flags, exp
    i, j, v, zero_based = co.load_graph(inp["habitat_file"][1], co.cfg_bool(cfg, "habitat_map_is_resistances", "True"))
    G, cc = co.network_graph(i, j, v)
    n = G.shape[0]

    def strengths(raw):
        raw = np.asarray(raw, dtype=np.float64).reshape(-1, 2).copy()
        if raw[:, 0].min() == 0 or zero_based:
            raw[:, 0] += 1
        return raw
    src = strengths(inp["source_file"][1]); gnd = strengths(inp["ground_file"][1])
    if fl["grnd_file_is_res"]:
        with np.errstate(divide="ignore"):
            gnd[:, 1] = 1.0 / gnd[:, 1]
    sources = np.zeros(n); grounds = np.zeros(n)
    sources[src[:, 0].astype(np.int64) - 1] = src[:, 1]
    grounds[gnd[:, 0].astype(np.int64) - 1] = gnd[:, 1]
    s, g, f = co.resolve_conflicts(sources, grounds, fl["policy"])
    return cb.AdvancedProblem(G, cc, s, g, f, None, None, None, solver), flags, exp


def check_advanced(r, exp, flags, map_tol=TOL):
    if flags.is_raster:
        if "curmap.asc" in exp:
            assert np.sum((r.curmap - exp["curmap.asc"]) ** 2) < map_tol
        if "voltmap.asc" in exp:
            assert np.sum((r.voltmap - exp["voltmap.asc"]) ** 2) < map_tol
        return
    x = exp["voltages.txt"].copy(); x[:, 0] += 1
    mine = np.column_stack([np.arange(1, len(r.voltages) + 1), r.voltages])
    assert np.all(np.abs(x - mine) <= np.sqrt(TOL))
    v = exp["node_currents.txt"].copy(); v[:, 0] += 1
    mine = np.column_stack([np.arange(1, len(r.voltages) + 1), r.node_currents])
    assert np.sum((sorted_rows(mine) - sorted_rows(v)) ** 2) < map_tol
    gr, gc, val = r.branch
    keep = ~np.isclose(val, 0.0, atol=1e-6)
    mine = np.column_stack([gr, gc, val])[keep]
    v = exp["branch_currents.txt"].copy(); v[:, :2] += 1
    assert mine.shape == v.shape
    assert np.sum((sorted_rows(mine) - sorted_rows(v)) ** 2) < map_tol


def onetoall_problem(golden, name):
    """-> (RasterData, flags, cfg, expected) for the one-to-all / all-to-one goldens."""
    cfg, inp, exp = co.load_case(golden, name)
    flags = cb.Flags.from_cfg(cfg)
    cellmap, polymap, meta, inc = co.load_raster_inputs(cfg, inp)
    pk = inp["point_file"]
    points_rc = co.read_point_map(pk[0], pk[1], meta)
    strengths = None
    if co.cfg_bool(cfg, "use_variable_source_strengths"):
        strengths = np.asarray(inp["variable_source_file"][1], dtype=np.float64).reshape(-1, 2).copy()
        if strengths[:, 0].min() == 0:
            strengths[:, 0] += 1
    data = cb.RasterData(cellmap, polymap, points_rc, strengths, inc)
    return data, flags, cfg, exp


def check_onetoall(r, exp, flags, map_tol=TOL):
    x = exp["resistances.out"]
    assert x.shape == r.resistances.shape
    assert np.all(np.abs(x - r.resistances) <= np.sqrt(TOL))             # test_utils.jl:125-128
    assert np.abs(x - r.resistances).max() <= 1e-6 * max(1.0, np.abs(x).max())
    o = flags.outputflags
    n = 0
    for pid, m in r.curmaps.items():                                      # only files the reference writes
        assert np.sum((m - exp[f"curmap_{pid}.asc"]) ** 2) < map_tol
        n += 1
    for pid, m in r.voltmaps.items():
        assert np.sum((m - exp[f"voltmap_{pid}.asc"]) ** 2) < map_tol
        n += 1
    if (o.write_cur_maps or o.write_cum_cur_map_only) and "cum_curmap.asc" in exp:
        assert np.sum((r.cum_curmap - exp["cum_curmap.asc"]) ** 2) < map_tol
    if o.write_max_cur_maps and "max_curmap.asc" in exp:
        assert np.sum((r.max_curmap - exp["max_curmap.asc"]) ** 2) < map_tol
    return n
