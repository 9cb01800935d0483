"""On-device raster assembly (circuitscape_b200/csrc/raster_assembly.cuh).

CPU: the kernels' index walk restated in Python (same slot order, same diagonal placement)
against the host assembly the rest of the suite uses -- pins the algorithm.
GPU: the CSR the device builds, downloaded through cs_b200_get_csr, against the same host
assembly, and a solve on the device-assembled handle against the host-assembled one."""
import numpy as np
import pytest
import scipy.sparse as sp

import circuitscape_b200 as cb
from circuitscape_b200 import graph


def host_laplacian(g, four, avg_res):
    nodemap = graph.construct_node_map(g, None)
    return graph.laplacian(graph.construct_graph(g, nodemap, avg_res, four)), nodemap


def emulate_device_walk(g, four, avg_res):
    """k_valid / scan / k_count / scan / k_fill of raster_assembly.cuh, one 'thread' per cell."""
    nrows, ncols = g.shape
    gf = np.asfortranarray(g).ravel(order="F")             # cell index = r + c * nrows
    valid = (gf > 0).astype(np.int64)
    nodeid = np.cumsum(valid) - valid
    n = int(nodeid[-1] + valid[-1])
    s2 = 1.4142135623730951

    def weight(a, b, diagonal):
        if avg_res:
            return 1.0 / (s2 * (1.0 / a + 1.0 / b) / 2.0) if diagonal else 1.0 / ((1.0 / a + 1.0 / b) / 2.0)
        return (a + b) / (2.0 * s2) if diagonal else (a + b) / 2.0

    rowcnt = np.zeros(n + 1, dtype=np.int64)
    for i in range(nrows * ncols):
        if not valid[i]:
            continue
        r, c = i % nrows, i // nrows
        cnt = 1
        for k in range(9):
            if k == 4:
                continue
            dr, dc = k % 3 - 1, k // 3 - 1
            if four and dr != 0 and dc != 0:
                continue
            rr, cc = r + dr, c + dc
            if rr < 0 or rr >= nrows or cc < 0 or cc >= ncols:
                continue
            cnt += valid[cc * nrows + rr]
        rowcnt[nodeid[i]] = cnt
    rowptr = np.cumsum(rowcnt) - rowcnt
    nnz = int(rowptr[n])
    colidx = np.full(nnz, -1, dtype=np.int64)
    vals = np.zeros(nnz)
    for i in range(nrows * ncols):
        if not valid[i]:
            continue
        r, c = i % nrows, i // nrows
        p = rowptr[nodeid[i]]
        deg = 0.0
        diag_pos = p
        for k in range(9):
            if k == 4:
                diag_pos = p
                p += 1
                continue
            dr, dc = k % 3 - 1, k // 3 - 1
            diagonal = dr != 0 and dc != 0
            if four and diagonal:
                continue
            rr, cc = r + dr, c + dc
            if rr < 0 or rr >= nrows or cc < 0 or cc >= ncols:
                continue
            j = cc * nrows + rr
            if not valid[j]:
                continue
            w = weight(gf[i], gf[j], diagonal)
            colidx[p] = nodeid[j]
            vals[p] = -w
            deg += w
            p += 1
        colidx[diag_pos] = nodeid[i]
        vals[diag_pos] = deg
    return sp.csr_matrix((vals, colidx, rowptr[: n + 1]), shape=(n, n))


def rasters():
    rng = np.random.default_rng(4)
    full = rng.uniform(0.2, 3.0, (9, 7))
    holes = rng.uniform(0.2, 3.0, (13, 11))
    holes[rng.random(holes.shape) < 0.25] = 0.0
    holes[0, 0] = -9999.0
    holes[5, 5] = np.nan
    thin = rng.uniform(0.5, 1.5, (1, 6))
    return {"full": full, "holes": holes, "thin": thin}


@pytest.mark.parametrize("four", [False, True])
@pytest.mark.parametrize("avg_res", [False, True])
@pytest.mark.parametrize("name", ["full", "holes", "thin"])
def test_device_walk_restated_on_cpu(name, four, avg_res):
    g = rasters()[name]
    gh = np.where(np.isnan(g) | (g <= 0), 0.0, g)
    L, _ = host_laplacian(gh, four, avg_res)
    E = emulate_device_walk(np.where(np.isnan(g), np.nan, g), four, avg_res)
    assert E.shape == L.shape
    E.sort_indices()
    # the device keeps explicit zeros nowhere and stores every diagonal (also 0 for isolated cells)
    D = (E - L).tocsr()
    assert np.abs(D.data).max(initial=0.0) < 1e-14 * np.abs(L.data).max()
    assert np.all(np.diff(E.indices)[np.diff(np.repeat(np.arange(E.shape[0]), np.diff(E.indptr))) == 0] > 0), \
        "column indices ascending within a row"


@pytest.mark.gpu
@pytest.mark.parametrize("four", [False, True])
@pytest.mark.parametrize("avg_res", [False, True])
@pytest.mark.parametrize("precision", ["double", "single"])
def test_device_assembly_matches_host(four, avg_res, precision):
    rng = np.random.default_rng(8)
    g = rng.uniform(0.2, 3.0, (157, 93))
    g[rng.random(g.shape) < 0.12] = 0.0
    g[3, 4] = -9999.0
    L, nodemap = host_laplacian(np.where(g > 0, g, 0.0), four, avg_res)
    solver = cb.CUDASolver(precision=precision, f32_compute=(precision == "single"), precond="jacobi")
    with cb.B200Factor.from_raster(g, solver, four_neighbors=four, avg_res=avg_res) as f:
        A = f.get_csr()
        assert f.n == L.shape[0] == int(nodemap.max())
    assert A.shape == L.shape
    rows = np.repeat(np.arange(A.shape[0]), np.diff(A.indptr))
    assert np.all(np.diff(A.indices)[np.diff(rows) == 0] > 0), "device column indices ascending within a row"
    assert np.all(np.diff(A.indptr) >= 1), "every node stores its diagonal"
    tol = 1e-14 if precision == "double" else 1e-6
    D = (A.astype(np.float64) - L).tocsr()
    assert np.abs(D.data).max(initial=0.0) <= tol * np.abs(L.data).max()
    A64 = A.astype(np.float64); A64.eliminate_zeros()
    L0 = L.copy(); L0.eliminate_zeros()
    assert A64.nnz == L0.nnz


@pytest.mark.gpu
@pytest.mark.parametrize("precond", ["jacobi", "amg"])
def test_solve_on_device_assembled_raster(precond):
    """pairs solved on the device-assembled operator == pairs solved on the uploaded one (the two
    matrices differ in the last bit -- different summation order of the diagonal -- so the PCG
    trajectories differ: agreement is at the solver tolerance, tightened here to 1e-10)"""
    rng = np.random.default_rng(9)
    g = rng.uniform(1.0, 10.0, (220, 180))
    g = 1.0 / g
    L, nodemap = host_laplacian(g, False, False)
    nodes = graph.focal_nodes(L.shape[0], 5, seed=7)
    src, dst = graph.all_pairs(nodes)
    with cb.B200Factor(L, cb.CUDASolver(precond=precond, rtol=1e-10)) as f0:
        r0 = f0.solve_pairs(src, dst, accumulate=True)
        c0, m0 = f0.read_currents()
    with cb.B200Factor.from_raster(g, cb.CUDASolver(precond=precond, rtol=1e-10)) as f1:
        r1 = f1.solve_pairs(src, dst, accumulate=True)
        c1, m1 = f1.read_currents()
    assert np.abs(r1["R"] - r0["R"]).max() <= 1e-9 * np.abs(r0["R"]).max()
    assert np.abs(c1 - c0).max() <= 1e-6 * np.abs(c0).max()


# ---- short-circuit polygons on the device (round 2) ------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("four,avg_res", [(False, False), (True, True), (False, True)])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_device_assembly_with_polygons_matches_host(seed, four, avg_res):
    """construct_node_map + construct_graph + laplacian! with a polygon map (src/raster/pairwise.jl:283-314):
    polygons over NODATA cells, a polygon without any valid cell, touching polygons, merged parallel
    adjacencies -- node map and Laplacian against the host assembly (circuitscape_b200/graph.py)."""
    import circuitscape_b200 as cb
    from circuitscape_b200 import graph
    rng = np.random.default_rng(seed)
    nr, nc = 41 + 7 * seed, 37
    g = rng.uniform(0.1, 1.0, size=(nr, nc))
    g[rng.random(g.shape) < 0.15] = 0.0                       # NODATA
    poly = np.zeros((nr, nc), dtype=np.int64)
    poly[3:9, 4:11] = 1
    poly[9:12, 4:8] = 2                                       # touches polygon 1
    poly[20:23, 20:30] = 7
    poly[30:33, 1:4] = 5
    g[30:33, 1:4] = 0.0                                       # polygon 5 has no valid cell at all
    poly[rng.integers(0, nr, 25), rng.integers(0, nc, 25)] = 9   # scattered cells of one polygon
    nm_host = graph.construct_node_map(g, poly)
    L_host = graph.laplacian(graph.construct_graph(g, nm_host, avg_res, four)).tocsr()
    f, nm_dev = cb.B200Factor.from_raster_polygons(g, poly, cb.CUDASolver(precond="jacobi"), four_neighbors=four,
                                                   avg_res=avg_res)
    with f:
        L_dev = f.get_csr()
    assert np.array_equal(nm_dev, nm_host)
    assert L_dev.shape == L_host.shape
    d = (L_dev - L_host).tocsr()
    assert abs(d).max() <= 1e-13 * abs(L_host).max()
    assert np.abs(np.asarray(L_dev.sum(axis=1))).max() <= 1e-12     # a Laplacian: zero row sums


@pytest.mark.gpu
def test_device_assembly_with_polygons_solves_like_the_host_path():
    import circuitscape_b200 as cb
    from circuitscape_b200 import graph
    rng = np.random.default_rng(3)
    g = rng.uniform(0.1, 1.0, size=(150, 140))
    g[rng.random(g.shape) < 0.05] = 0.0
    poly = np.zeros(g.shape, dtype=np.int64)
    poly[10:30, 10:25] = 3
    poly[100:140, 60:70] = 4
    nm = graph.construct_node_map(g, poly)
    L = graph.laplacian(graph.construct_graph(g, nm, False, False))
    comp = max(graph.connected_components(L), key=len)
    a, b = int(comp[5]) - 1, int(comp[-7]) - 1
    with cb.B200Factor(L, cb.CUDASolver(rtol=1e-9)) as f0:
        R0 = f0.solve_pairs([a], [b])["R"][0]
    f, nm_dev = cb.B200Factor.from_raster_polygons(g, poly, cb.CUDASolver(rtol=1e-9))
    with f:
        R1 = f.solve_pairs([a], [b])["R"][0]
    assert np.array_equal(nm_dev, nm) and abs(R0 - R1) <= 1e-7 * R0
