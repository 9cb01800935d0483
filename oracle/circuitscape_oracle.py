"""CPU oracle for the Circuitscape.jl Laplacian-solve hot path.

TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  Only `tests/`, `__graft_entry__.smoke()`
and `bench.py`'s CPU-baseline legs may import this module.  The product path
(`circuitscape_b200/`) never does; it fails loudly without its CUDA library.

This is a from-scratch numpy/scipy restatement of what the reference computes
on the path  compute -> raster_pairwise/advanced | network_pairwise/advanced ->
single_ground_all_pairs / advanced_kernel -> solve_linear_system -> postprocess.
Every function cites the reference file:line (relative to /root/reference) it
follows.  Parity status: PINNED -- `tests/test_oracle_golden.py` checks it
against every golden vector the reference's own integration suite holds for
this path (test/output_verify sgVerify1-17, sgNetworkVerify1-3, mgVerify1-6,
mgNetworkVerify1-3, oneToAllVerify1-13, allToOneVerify1-12; packed by tests/golden/make_fixtures.py) with the
reference's own tolerances (test/test_utils.jl:72-73,147,196,217-226).

Linear solves: the reference's arithmetic lives in un-vendored Julia packages
(Krylov.jl 0.10 `cg`, AlgebraicMultigrid.jl 1.2 `smoothed_aggregation`,
SuiteSparse CHOLMOD).  Those are restated by their published algorithms:
  * "direct"  : exact sparse LU of the grounded system (ground truth; the
                unique solution every reference solver approximates),
  * "cholmod" : LU of  A + 10*eps*I  (core.jl:519-523) -- the reference's
                direct-solver regularisation,
  * "cg+amg"  : smoothed-aggregation AMG (Gauss-Seidel pre/post, pseudo-inverse
                coarse solve) preconditioned CG, rtol 1e-6 (core.jl:161-167,639)
                -- see oracle/amg.py.
Conventions: node ids are 1-based with 0 = "no node" exactly like the
reference's nodemap; matrices are scipy CSR/CSC with 0-based row = node-1.
"""
from __future__ import annotations

import json
from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp
import scipy.sparse.csgraph as csgraph
import scipy.sparse.linalg as spla

NODATA = -9999.0
RESISTANCE_INVALID = -777.0  # consts.jl:45
TRUE = ("True", "true", "1")  # config.jl:55-57


# ----------------------------------------------------------------------------
# config (config.jl:87-135)
# ----------------------------------------------------------------------------
def cfg_bool(cfg, key, default="false"):
    return cfg.get(key, default) in TRUE


def cfg_flags(cfg):
    """OutputFlags + raster flags (raster/pairwise.jl:32-52, utils.jl:33-38)."""
    return dict(
        write_volt_maps=cfg_bool(cfg, "write_volt_maps"),
        write_cur_maps=cfg_bool(cfg, "write_cur_maps"),
        write_cum_cur_map_only=cfg_bool(cfg, "write_cum_cur_map_only"),
        write_max_cur_maps=cfg_bool(cfg, "write_max_cur_maps"),
        set_null_currents_to_nodata=cfg_bool(cfg, "set_null_currents_to_nodata"),
        set_null_voltages_to_nodata=cfg_bool(cfg, "set_null_voltages_to_nodata"),
        log_transform_maps=cfg_bool(cfg, "log_transform_maps"),
        four_neighbors=cfg_bool(cfg, "connect_four_neighbors_only"),
        avg_res=cfg_bool(cfg, "connect_using_avg_resistances"),
        policy=cfg.get("remove_src_or_gnd", "keepall"),
        grnd_file_is_res=cfg_bool(cfg, "ground_file_is_resistances", "True"),
    )


# ----------------------------------------------------------------------------
# input conventions (io.jl)
# ----------------------------------------------------------------------------
def colmajor_nonzero(mask):
    """Julia `findall` on a Matrix walks column-major; return (rows, cols) 0-based."""
    c, r = np.nonzero(np.asarray(mask).T)
    return r, c


def read_cellmap(raw, is_res):
    """io.jl:91-111: resistance -> conductance, NODATA(-9999) -> 0."""
    raw = np.asarray(raw, dtype=np.float64)
    nod = raw == NODATA
    if is_res:
        if np.any(raw == 0):
            raise ValueError("zero resistance values are not currently supported for habitat maps")
        with np.errstate(divide="ignore"):
            g = 1.0 / raw
    else:
        g = raw.copy()
    g[nod] = 0.0
    return g


def read_polymap(raw, nodata_as=0):
    """io.jl:160-192: NODATA -> nodata_as (unless nodata_as == -1)."""
    p = np.asarray(raw, dtype=np.float64).copy()
    if nodata_as != -1:
        p[p == NODATA] = nodata_as
    return p


def apply_mask(cellmap, mask_raw):
    """io.jl:511-515 (update!)."""
    m = read_polymap(mask_raw)
    m = np.where(m > 0, 1.0, 0.0)
    out = cellmap * m
    if out.sum() == 0:
        raise ValueError("Mask file deleted everything!")
    return out


def read_point_map(kind, raw, meta):
    """io.jl:194-253.  Returns (i, j, v) 1-based row, col and id, sorted by id."""
    ncols, nrows, xll, yll, cs = [float(x) for x in meta]
    if kind == "txtlist":
        raw = np.asarray(raw, dtype=np.float64)
        I, J, v = raw[:, 1], raw[:, 2], raw[:, 0]
        i = np.ceil(nrows - (J - yll) / cs).astype(np.int64)
        j = np.ceil((I - xll) / cs).astype(np.int64)
        v = v.astype(np.int64)
    else:
        pm = read_polymap(raw)
        r, c = colmajor_nonzero(pm != 0)
        i, j, v = r + 1, c + 1, pm[r, c].astype(np.int64)
    i, j, v = list(i), list(j), list(v)
    # io.jl:222-229 deletes at indices computed before any deletion (stale after
    # the first one); restated as-is.
    neg = [k for k, x in enumerate(v) if x < 0]
    for index in neg:
        if index < len(v):
            del i[index], j[index], v[index]
    i, j, v = np.array(i, dtype=np.int64), np.array(j, dtype=np.int64), np.array(v, dtype=np.int64)
    order = np.argsort(v, kind="stable")
    i, j, v = i[order], j[order], v[order]
    if i.min() < 0 or j.min() < 0 or i.max() > nrows or j.max() > ncols:
        raise ValueError("At least one focal node location falls outside of habitat map")
    if len(np.unique(v)) < 2:
        raise ValueError("Less than two valid focal nodes found.")
    return i, j, v


@dataclass
class IncludePairs:
    mode: str
    point_ids: np.ndarray
    mat: np.ndarray


def read_included_pairs(kind, raw, meta):
    """io.jl:328-385."""
    raw = np.asarray(raw, dtype=np.float64)
    if kind == "pairs_aagrid":
        minval, maxval = float(meta[0]), float(meta[1])
        point_ids = raw[1:, 0].astype(np.int64)
        m = raw[1:, 1:].copy()
        m[m > maxval] = 0
        return IncludePairs("include", point_ids, (m >= minval).astype(np.int64))
    mode = kind.split("_")[-1]
    ids = np.unique(raw).astype(np.int64)
    ids = ids[ids != 0]
    mat = np.zeros((len(ids), len(ids)), dtype=np.int64)
    pos = {int(p): k for k, p in enumerate(ids)}
    for a, b in raw:
        if int(a) in pos and int(b) in pos:
            mat[pos[int(a)], pos[int(b)]] = 1
            mat[pos[int(b)], pos[int(a)]] = 1
    return IncludePairs(mode, ids, mat)


def generate_exclude_pairs(points_rc, inc):
    """raster/pairwise.jl:240-269.  Returns (pruned points_rc, exclude set)."""
    ex = set()
    ids, mat = inc.point_ids, inc.mat
    if inc.mode == "include":
        keep = np.isin(points_rc[2], ids)  # prune_points!  raster/onetoall.jl:169-180
        points_rc = tuple(a[keep] for a in points_rc)
        for j in range(mat.shape[1]):
            for i in range(mat.shape[0]):
                if mat[i, j] == 0 and mat[j, i] == 0:
                    ex.add((int(ids[i]), int(ids[j])))
    else:
        for j in range(mat.shape[1]):
            for i in range(mat.shape[0]):
                if mat[i, j] == 1 and mat[j, i] == 1:
                    ex.add((int(ids[i]), int(ids[j])))
    return points_rc, ex


# ----------------------------------------------------------------------------
# graph assembly (raster/pairwise.jl:271-367, core.jl:608-634)
# ----------------------------------------------------------------------------
def relabel(nodemap, offset=0):
    """raster/pairwise.jl:303-314: compact labels, preserving order."""
    nz = nodemap != 0
    if not nz.any():
        return nodemap
    _, inv = np.unique(nodemap[nz], return_inverse=True)
    nodemap[nz] = inv + offset
    return nodemap


def construct_node_map(gmap, polymap):
    """raster/pairwise.jl:271-301: column-major numbering of cells with g > 0;
    every cell of a short-circuit polygon (even NODATA ones) takes the node of the
    polygon's first valid cell; labels compacted."""
    gmap = np.asarray(gmap)
    nodemap = np.zeros(gmap.shape, dtype=np.int64)
    ind = gmap > 0
    r, c = colmajor_nonzero(ind)
    nodemap[r, c] = np.arange(1, len(r) + 1)
    if polymap is None or np.size(polymap) == 0:
        return nodemap
    polymap = np.asarray(polymap)
    pruned = np.where(ind, polymap, 0)
    for polynum in np.unique(polymap):
        if polynum == 0:
            continue
        r1, c1 = colmajor_nonzero(pruned == polynum)
        if len(r1) > 0:
            nodemap[polymap == polynum] = nodemap[r1[0], c1[0]]
    return relabel(nodemap, 1)


def _avg_fns(avg_res):
    """raster/pairwise.jl:364-367 (arguments are conductances)."""
    s2 = np.sqrt(2.0)
    with np.errstate(divide="ignore"):
        if avg_res:
            return (lambda x, y: 1.0 / ((1.0 / x + 1.0 / y) / 2.0),
                    lambda x, y: 1.0 / (s2 * (1.0 / x + 1.0 / y) / 2.0))
        return (lambda x, y: (x + y) / 2.0, lambda x, y: (x + y) / (2.0 * s2))


def construct_graph(gmap, nodemap, avg_res, four_neighbors):
    """raster/pairwise.jl:316-362.  E, S, SE, NE edges from every node cell, summed
    over duplicates (merged polygon nodes), then symmetrised a + a'.  Vectorised
    over cells; the order of summation of duplicates differs from the reference's
    `sparse(I,J,V)` only in fp round-off of parallel conductances."""
    gmap = np.asarray(gmap, dtype=np.float64)
    f1, f2 = _avg_fns(avg_res)
    nr, nc = gmap.shape
    nm = nodemap
    I, J, V = [], [], []

    def emit(a_sl, b_sl, f):
        na, nb = nm[a_sl], nm[b_sl]
        ok = (na != 0) & (nb != 0)
        with np.errstate(divide="ignore", invalid="ignore"):
            val = f(gmap[a_sl], gmap[b_sl])
        I.append(na[ok]); J.append(nb[ok]); V.append(val[ok])

    emit((slice(None), slice(0, nc - 1)), (slice(None), slice(1, nc)), f1)          # east
    emit((slice(0, nr - 1), slice(None)), (slice(1, nr), slice(None)), f1)          # south
    if not four_neighbors:
        emit((slice(0, nr - 1), slice(0, nc - 1)), (slice(1, nr), slice(1, nc)), f2)  # south-east
        emit((slice(1, nr), slice(0, nc - 1)), (slice(0, nr - 1), slice(1, nc)), f2)  # north-east
    I = np.concatenate(I) - 1
    J = np.concatenate(J) - 1
    V = np.concatenate(V)
    m = int(nm.max())
    a = sp.coo_matrix((V, (I, J)), shape=(m, m)).tocsr()
    return (a + a.T).tocsr()


def laplacian(a):
    """core.jl:608-634: off-diagonals -> -g_ij, diagonal -> sum_j g_ij (stored
    diagonal entries of the adjacency are discarded)."""
    a = sp.csr_matrix(a, dtype=np.float64)
    a = a - sp.diags(a.diagonal())
    deg = np.asarray(a.sum(axis=1)).ravel()
    return (sp.diags(deg) - a).tocsr()


def connected_components(G):
    """Graphs.connected_components(SimpleGraph(G)) (raster/pairwise.jl:171,214):
    edges where the stored value != 0; components ordered by their smallest
    node, nodes ascending.  Returns list of 1-based node-id arrays."""
    A = sp.csr_matrix(G).copy()
    A.data = (A.data != 0).astype(np.int8)
    A.eliminate_zeros()
    ncomp, lab = csgraph.connected_components(A, directed=False)
    order = np.argsort(lab, kind="stable")
    counts = np.bincount(lab, minlength=ncomp)
    comps = np.split(order + 1, np.cumsum(counts)[:-1])
    comps.sort(key=lambda c: c[0])
    return comps


# ----------------------------------------------------------------------------
# linear solves (core.jl:519-523, 636-653)
# ----------------------------------------------------------------------------
class ResidualError(RuntimeError):
    pass


def _check_residual(A, x, b, who):
    """core.jl:640-641,648-651: true relative residual gate 1e-4."""
    for col in range(b.shape[1]):
        nb = np.linalg.norm(b[:, col])
        res = np.linalg.norm(A @ x[:, col] - b[:, col]) / nb
        if not res < 1e-4:
            raise ResidualError(f"{who} solver residual {res} exceeds tolerance 1e-4 for column {col + 1}")


def solve_pairs_direct(A, src, dst):
    """Ground truth for the pairwise system  A v = e_dst - e_src  on a connected
    component: pin v[src] = 0 (delete row/col src), solve the SPD reduced system
    exactly.  Identical to the reference's `v .- v[src]` (core.jl:231,466-472) of
    any solution of the singular system.  Batches pairs sharing a source."""
    A = sp.csc_matrix(A)
    n = A.shape[0]
    out = np.zeros((n, len(src)))
    src = np.asarray(src); dst = np.asarray(dst)
    for s in np.unique(src):
        cols = np.nonzero(src == s)[0]
        keep = np.ones(n, dtype=bool); keep[s] = False
        idx = np.nonzero(keep)[0]
        lu = spla.splu(A[idx][:, idx].tocsc())
        pos = np.full(n, -1); pos[idx] = np.arange(n - 1)
        rhs = np.zeros((n - 1, len(cols)))
        rhs[pos[dst[cols]], np.arange(len(cols))] = 1.0
        sol = lu.solve(rhs)
        out[np.ix_(idx, cols)] = sol.reshape(n - 1, len(cols))
    return out


def solve_cholmod_like(A, rhs):
    """core.jl:519-523 + 646-653: factor (A + 10 eps I), solve, gate residual."""
    A = sp.csc_matrix(A)
    n = A.shape[0]
    lu = spla.splu((A + 10 * np.finfo(np.float64).eps * sp.identity(n)).tocsc())
    rhs2 = rhs.reshape(n, -1)
    x = lu.solve(rhs2).reshape(n, -1)
    _check_residual(A, x, rhs2, "CHOLMOD")
    return x.reshape(rhs.shape)


def make_pair_solver(kind):
    """Returns f(A, src, dst) -> voltages (n x k), already shifted so v[src] = 0."""
    if kind == "direct":
        return solve_pairs_direct
    if kind == "cholmod":
        def f(A, src, dst):
            n = A.shape[0]
            rhs = np.zeros((n, len(src)))
            rhs[src, np.arange(len(src))] = -1.0  # core.jl:224-226,459-460
            rhs[dst, np.arange(len(src))] = 1.0
            v = solve_cholmod_like(A, rhs)
            return v - v[src, np.arange(len(src))][None, :]  # core.jl:231,466-472
        return f
    if kind == "cg+amg":
        from . import amg

        def f(A, src, dst):
            n = A.shape[0]
            A = sp.csr_matrix(A, dtype=np.float64).copy()
            A.data = A.data + np.finfo(np.float64).eps * np.linalg.norm(A.data)  # core.jl:161
            ml = amg.smoothed_aggregation(A)
            out = np.zeros((n, len(src)))
            for c, (s, d) in enumerate(zip(src, dst)):
                b = np.zeros(n); b[s] = -1.0; b[d] = 1.0
                v, _ = amg.pcg(A, b, ml, rtol=1e-6, itmax=100_000)  # core.jl:639
                _check_residual(A, v[:, None], b[:, None], "CG")
                out[:, c] = v - v[s]
            return out
        return f
    raise ValueError(kind)


# ----------------------------------------------------------------------------
# currents (out.jl:150-303)
# ----------------------------------------------------------------------------
def branch_currents_posneg(G, v, pos):
    """out.jl:250-290.  For every stored upper-triangular entry (row < col) in
    column-major (col, then row) order:  |G_rc| (v_r - v_c)  (pos) or negated;
    entries with |b / max(b)| < 1e-8 are zeroed.  Returns (rows, cols, b)."""
    Gc = sp.csc_matrix(G)
    Gc.sort_indices()
    col = np.repeat(np.arange(Gc.shape[1]), np.diff(Gc.indptr))
    row = Gc.indices
    up = col > row
    r, c, val = row[up], col[up], Gc.data[up]
    d = (v[r] - v[c]) if pos else (v[c] - v[r])
    b = np.abs(val) * d
    if len(b):
        maxcur = b.max()
        with np.errstate(divide="ignore", invalid="ignore"):
            b = np.where(np.abs(b / maxcur) < 1e-8, 0.0, b)
    return r, c, b


def node_currents_posneg(G, v, finitegrounds, pos):
    """out.jl:186-207: antisymmetrise B - B', drop negatives, add finite-ground
    currents on the diagonal, column sums."""
    n = G.shape[0]
    r, c, b = branch_currents_posneg(G, v, pos)
    # (B - B')[r,c] = b, [c,r] = -b ; keep positives ; column sum
    s = np.zeros(n)
    np.add.at(s, c, np.where(b > 0, b, 0.0))
    np.add.at(s, r, np.where(-b > 0, -b, 0.0))
    if finitegrounds is not None and not (len(finitegrounds) >= 1 and finitegrounds[0] == NODATA):
        fg = finitegrounds * v
        fg = np.where(fg < 0, -fg, 0.0) if pos else np.where(fg > 0, fg, 0.0)
        s = s + fg
    return s


def get_node_currents(G, v, finitegrounds=None):
    """out.jl:178-184: elementwise max of the pos and neg passes."""
    p = node_currents_posneg(G, v, finitegrounds, True)
    q = node_currents_posneg(G, v, finitegrounds, False)
    return np.where(p > q, p, q)


def get_branch_currents_abs(G, v):
    """out.jl:154-158 (network mode): abs of the `pos` branch currents, as
    (row, col, value) with row < col in CSC order of the upper triangle."""
    r, c, b = branch_currents_posneg(G, v, True)
    return r, c, np.abs(b)


def scatter_to_raster(values, local_nodemap):
    """out.jl:160-171, 421-434: raster[i,j] = values[local_nodemap[i,j]] or 0."""
    out = np.zeros(local_nodemap.shape)
    nz = local_nodemap != 0
    out[nz] = values[local_nodemap[nz] - 1]
    return out


def construct_local_node_map(nodemap, comp, polymap):
    """utils.jl:10-30."""
    local = np.zeros_like(nodemap)
    idx = np.isin(nodemap, comp)
    local[idx] = nodemap[idx]
    if np.array_equal(nodemap, local):
        return local
    if polymap is None or np.size(polymap) == 0:
        r, c = colmajor_nonzero(local != 0)
        local[r, c] = np.arange(1, len(r) + 1)
        return local
    lp = np.zeros_like(local)
    lp[idx] = polymap[idx]
    return construct_node_map(local, lp)


def process_grid(cmap, cellmap, log_transform, set_null_to_nodata):
    """out.jl:305-319."""
    if log_transform:
        with np.errstate(divide="ignore", invalid="ignore"):
            cmap = np.where(cmap > 0, np.log10(np.where(cmap > 0, cmap, 1.0)), NODATA)
    if set_null_to_nodata:
        cmap = np.where(cellmap == 0, NODATA, cmap)
    return cmap


# ----------------------------------------------------------------------------
# pairwise driver (core.jl:96-305, 312-515, 537-603, 685-739)
# ----------------------------------------------------------------------------
@dataclass
class GraphProblem:
    """core.jl:10-22."""
    G: sp.csr_matrix
    cc: list
    points: np.ndarray        # graph node (1-based, 0 = none) per focal point
    user_points: np.ndarray   # user ids
    exclude_pairs: set
    nodemap: np.ndarray | None = None
    polymap: np.ndarray | None = None
    cellmap: np.ndarray | None = None
    is_raster: bool = True
    coords: tuple | None = None  # network: (i, j) 1-based edge list for cum branch currents


@dataclass
class PairwiseResult:
    resistances: np.ndarray                 # (P+1)x(P+1) with ids (core.jl:294-299)
    voltmaps: dict = field(default_factory=dict)    # (id_i,id_j) -> raster | (nodes, volts)
    curmaps: dict = field(default_factory=dict)     # (id_i,id_j) -> raster | node currents
    branch: dict = field(default_factory=dict)      # network: (id_i,id_j) -> (r, c, |b|) 1-based global
    cum_curmap: np.ndarray | None = None
    max_curmap: np.ndarray | None = None
    cum_node: np.ndarray | None = None
    cum_branch: np.ndarray | None = None
    num_solves: int = 0


def enumerate_pairs(prob, comp, shortcut):
    """core.jl:148-250 / 386-424: per component, unordered pairs of *unique graph
    nodes* holding focal points, each fanned out to every (c_i, c_j) combination
    of focal indices on those nodes that is not excluded.
    Returns (csub, list of (src_node, dst_node, [(c_i, c_j), ...]), zero_pairs)."""
    pts, ids, ex = prob.points, prob.user_points, prob.exclude_pairs
    comp_set = set(int(x) for x in comp)
    csub = []
    for x in pts:
        if int(x) in comp_set and int(x) not in csub:
            csub.append(int(x))
    solves, zero = [], []
    npts = 1 if shortcut else len(csub)
    for pi in range(min(npts, len(csub))):
        s = csub[pi]
        si = [k for k, x in enumerate(pts) if x == s]
        for a in range(len(si)):            # smash_repeats!  core.jl:588-603
            for b in range(a + 1, len(si)):
                zero.append((si[a], si[b]))
        for pj in range(pi + 1, len(csub)):
            d = csub[pj]
            di = [k for k, x in enumerate(pts) if x == d]
            fan = [(ci, cj) for ci in si for cj in di if (int(ids[ci]), int(ids[cj])) not in ex]
            if fan:
                solves.append((s, d, fan))
    return csub, solves, zero


def single_ground_all_pairs(prob, flags, solver="direct", cellmap_for_null=None):
    """core.jl:70-72 -> solve (core.jl:96-305).  Returns PairwiseResult."""
    pair_solver = make_pair_solver(solver) if isinstance(solver, str) else solver
    P = len(prob.points)
    R = -np.ones((P, P))
    res = PairwiseResult(resistances=None)
    want_maps = (flags["write_volt_maps"] or flags["write_cur_maps"] or
                 flags["write_cum_cur_map_only"] or flags["write_max_cur_maps"])
    shortcut = prob.is_raster and not want_maps and not prob.exclude_pairs  # core.jl:137-145
    voltmatrix = np.zeros((P, P))
    shortcut_res = -np.ones((P, P))
    if prob.is_raster:
        res.cum_curmap = np.zeros(prob.cellmap.shape)                       # utils.jl:122-130
        res.max_curmap = np.full(prob.cellmap.shape, NODATA) if flags["write_max_cur_maps"] else None
    else:
        res.cum_node = np.zeros(prob.G.shape[0])                              # utils.jl:132-142
        res.cum_branch = np.zeros(len(prob.coords[0]))
        coord_pos = {}
        for k, (a, b) in enumerate(zip(prob.coords[0], prob.coords[1])):
            coord_pos.setdefault((int(a), int(b)), k)

    Gcsr = sp.csr_matrix(prob.G)
    for comp in prob.cc:
        csub, solves, zero = enumerate_pairs(prob, comp, shortcut)
        if not csub:
            continue
        for a, b in zero:
            R[a, b] = R[b, a] = 0.0
        idx = np.asarray(comp) - 1
        A = Gcsr[idx][:, idx].tocsr()
        pos = {int(node): k for k, node in enumerate(comp)}
        local_nodemap = None
        if prob.is_raster and not shortcut:
            local_nodemap = construct_local_node_map(prob.nodemap, comp, prob.polymap)
        if solves:
            src = np.array([pos[s] for s, _, _ in solves])
            dst = np.array([pos[d] for _, d, _ in solves])
            V = pair_solver(A, src, dst)
            res.num_solves += len(solves)
        for col, (s, d, fan) in enumerate(solves):
            v = V[:, col]
            r = v[dst[col]] - v[src[col]]                                    # core.jl:232
            node_cur = None
            for ci, cj in fan:
                R[ci, cj] = R[cj, ci] = r
                key = (int(prob.user_points[ci]), int(prob.user_points[cj]))
                if shortcut:                                                 # core.jl:685-703
                    for i in range(1, P):
                        p = int(prob.points[i])
                        if p in pos:
                            voltmatrix[i, cj] = 1.0 - v[pos[p]] / r
                    continue
                if prob.is_raster:
                    if flags["write_volt_maps"]:
                        vm = scatter_to_raster(v, local_nodemap)
                        res.voltmaps[key] = process_grid(vm, prob.cellmap, False,
                                                         flags["set_null_voltages_to_nodata"])
                    if node_cur is None:
                        node_cur = get_node_currents(A, v)
                    cmap = scatter_to_raster(node_cur, local_nodemap)
                    cmap = process_grid(cmap, prob.cellmap, flags["log_transform_maps"],
                                        flags["set_null_currents_to_nodata"])
                    res.cum_curmap += cmap                                   # out.jl:100-107
                    if res.max_curmap is not None:
                        res.max_curmap = np.maximum(res.max_curmap, cmap)
                    if flags["write_cur_maps"] and not flags["write_cum_cur_map_only"]:
                        res.curmaps[key] = cmap
                else:
                    if flags["write_volt_maps"]:
                        res.voltmaps[key] = (np.asarray(comp), v.copy())
                    if node_cur is None:
                        node_cur = get_node_currents(A, v)
                        br = get_branch_currents_abs(A, v)
                    gr, gc = np.asarray(comp)[br[0]], np.asarray(comp)[br[1]]
                    for a_, b_, val in zip(gr, gc, br[2]):                   # out.jl:65-76
                        k = coord_pos.get((int(a_), int(b_)))
                        if k is None:
                            k = coord_pos.get((int(b_), int(a_)))
                        res.cum_branch[k] += val
                    res.cum_node[np.asarray(comp) - 1] += node_cur           # out.jl:78-83
                    res.curmaps[key] = (np.asarray(comp), node_cur)
                    res.branch[key] = (gr, gc, br[2])
        if shortcut:
            anchor = next(k for k, x in enumerate(prob.points) if x == csub[0])
            _update_shortcut(anchor, voltmatrix, shortcut_res, R, prob.points, comp)
    if shortcut:
        R = shortcut_res
    np.fill_diagonal(R, 0.0)
    out = np.zeros((P + 1, P + 1))
    out[0, 1:] = prob.user_points
    out[1:, 0] = prob.user_points
    out[1:, 1:] = R
    res.resistances = out
    if prob.is_raster:
        res.cum_curmap = np.where(res.cum_curmap < NODATA, NODATA, res.cum_curmap)   # utils.jl:114-120
        if res.max_curmap is not None:
            res.max_curmap = np.where(res.max_curmap < NODATA, NODATA, res.max_curmap)
    return res


def _update_shortcut(anchor, voltmatrix, shortcut, resistances, points, comp):
    """core.jl:706-739 (R_2x = 2 R_12 V_x2 + R_1x - R_12)."""
    comp_set = set(int(x) for x in comp)
    check = [int(p) in comp_set for p in points]
    l = resistances.shape[0]
    for px in range(l):
        if not check[px]:
            continue
        R1x = resistances[anchor, px]
        if R1x == -1:
            continue
        shortcut[px, anchor] = shortcut[anchor, px] = R1x
        for p2 in range(px, l):
            if not check[p2]:
                continue
            R12 = resistances[anchor, p2]
            if R12 == -1:
                continue
            if R1x != RESISTANCE_INVALID:
                shortcut[anchor, p2] = shortcut[p2, anchor] = R12
                Vx = voltmatrix[px, p2]
                R2x = 2 * R12 * Vx + R1x - R12
                if shortcut[p2, px] != RESISTANCE_INVALID:
                    shortcut[p2, px] = shortcut[px, p2] = R2x
            else:
                shortcut[px, :] = RESISTANCE_INVALID
                shortcut[:, px] = RESISTANCE_INVALID


def compute_3col(r):
    """out.jl:12-26."""
    fp = r[1:, 0]
    l = len(fp)
    rows = []
    for i in range(l):
        for j in range(i + 1, l):
            rows.append((fp[i], fp[j], r[j + 1, i + 1]))
    return np.array(rows).reshape(-1, 3)


# ----------------------------------------------------------------------------
# raster pairwise front end (raster/pairwise.jl:14-135, 192-238, 369-442)
# ----------------------------------------------------------------------------
def load_raster_inputs(cfg, inputs):
    """io.jl:420-508 for in-memory arrays.  `inputs[key]` = (kind, raw, meta)."""
    hab = inputs["habitat_file"]
    cellmap = read_cellmap(hab[1], cfg_bool(cfg, "habitat_map_is_resistances", "True"))
    meta = hab[2]
    polymap = read_polymap(inputs["polygon_file"][1]).astype(np.int64) if cfg_bool(cfg, "use_polygons") else None
    if cfg_bool(cfg, "use_mask"):
        cellmap = apply_mask(cellmap, inputs["mask_file"][1])
    inc = None
    if cfg_bool(cfg, "use_included_pairs"):
        k = inputs["included_pairs_file"]
        inc = read_included_pairs(k[0], k[1], k[2])
    return cellmap, polymap, meta, inc


def raster_pairwise(cfg, inputs, solver="direct"):
    """raster/pairwise.jl:14-30."""
    flags = cfg_flags(cfg)
    cellmap, polymap, meta, inc = load_raster_inputs(cfg, inputs)
    pk = inputs["point_file"]
    points_rc = read_point_map(pk[0], pk[1], meta)
    if len(points_rc[0]) != len(np.unique(points_rc[2])):
        return _pt_file_polygons_path(cellmap, polymap, points_rc, inc, flags, solver)
    exclude = set()
    if inc is not None:
        points_rc, exclude = generate_exclude_pairs(points_rc, inc)
    nodemap = construct_node_map(cellmap, polymap)
    G = laplacian(construct_graph(cellmap, nodemap, flags["avg_res"], flags["four_neighbors"]))
    cc = connected_components(G)
    points = nodemap[points_rc[0] - 1, points_rc[1] - 1]
    prob = GraphProblem(G, cc, points, points_rc[2], exclude, nodemap, polymap, cellmap, True)
    return single_ground_all_pairs(prob, flags, solver)


def create_new_polymap(gmap, polymap, points_rc, pt1, pt2):
    """raster/pairwise.jl:369-442 (pairwise branch: point_map empty)."""
    rr, cc_, ids = points_rc
    if polymap is None or np.size(polymap) == 0:
        newpoly = np.zeros(gmap.shape, dtype=np.int64)
        for p in (pt1, pt2):
            sel = ids == p
            newpoly[rr[sel] - 1, cc_[sel] - 1] = p
        return newpoly
    newpoly = polymap.copy()
    k = polymap.max()
    for p in (pt1, pt2):
        idx = np.nonzero(ids == p)[0]
        if len(idx) == 1:
            continue
        vals_at = polymap[rr[idx] - 1, cc_[idx] - 1]
        if np.all(vals_at == 0):
            newpoly[rr[idx] - 1, cc_[idx] - 1] = k + 1
            k += 1
        else:
            nz = idx[vals_at != 0]
            if len(nz) == 1:
                # reference line 428 reads an undefined variable (`overlap`) here and
                # would throw; no golden case reaches it.
                raise NotImplementedError("reference raises UndefVarError on this branch")
            vals = polymap[rr[nz] - 1, cc_[nz] - 1]
            newpoly[np.isin(polymap, vals)] = k + 1
            k += 1
    return newpoly


def _pt_file_polygons_path(cellmap, polymap, points_rc, inc, flags, solver):
    """raster/pairwise.jl:72-135: focal *regions*: graph rebuilt per pair."""
    exclude = set()
    if inc is not None:
        points_rc, exclude = generate_exclude_pairs(points_rc, inc)
    pts = []
    for p in points_rc[2]:
        if int(p) not in pts:
            pts.append(int(p))
    n = len(pts)
    R = -np.ones((n, n))
    out = PairwiseResult(resistances=None)
    out.cum_curmap = np.zeros(cellmap.shape)
    out.max_curmap = np.full(cellmap.shape, NODATA) if flags["write_max_cur_maps"] else None
    for i in range(n):
        for j in range(i + 1, n):
            pt1, pt2 = pts[i], pts[j]
            if (pt1, pt2) in exclude or (pt2, pt1) in exclude:
                continue
            newpoly = create_new_polymap(cellmap, polymap, points_rc, pt1, pt2)
            nodemap = construct_node_map(cellmap, newpoly)
            a = construct_graph(cellmap, nodemap, flags["avg_res"], flags["four_neighbors"])
            G = laplacian(a)
            cc = connected_components(G)
            x = int(np.nonzero(points_rc[2] == pt1)[0][0])
            y = int(np.nonzero(points_rc[2] == pt2)[0][0])
            c1 = nodemap[points_rc[0][x] - 1, points_rc[1][x] - 1]
            c2 = nodemap[points_rc[0][y] - 1, points_rc[1][y] - 1]
            prob = GraphProblem(G, cc, np.array([c1, c2]), np.array([pt1, pt2]), set(),
                                nodemap, newpoly, cellmap, True)
            r = single_ground_all_pairs(prob, flags, solver)
            R[i, j] = R[j, i] = r.resistances[1, 2]
            out.num_solves += r.num_solves
            out.voltmaps.update(r.voltmaps)
            out.curmaps.update(r.curmaps)
            out.cum_curmap += r.cum_curmap   # shared `cum` object in the reference
            if out.max_curmap is not None:
                out.max_curmap = np.maximum(out.max_curmap, r.max_curmap)
    np.fill_diagonal(R, 0.0)
    full = np.zeros((n + 1, n + 1))
    full[0, 1:] = pts
    full[1:, 0] = pts
    full[1:, 1:] = R
    out.resistances = full
    return out


# ----------------------------------------------------------------------------
# network front ends (network/pairwise.jl, network/advanced.jl, io.jl:49-89,387-418)
# ----------------------------------------------------------------------------
def load_graph(raw, is_res):
    """io.jl:49-72 + 401-403: 0-based files are shifted to 1-based."""
    raw = np.asarray(raw, dtype=np.float64)
    i, j, v = raw[:, 0].astype(np.int64), raw[:, 1].astype(np.int64), raw[:, 2].copy()
    mn = min(i.min(), j.min())
    if mn > 1:
        raise ValueError("resistance file must start counting nodes from 1 (or 0)")
    zero_based = mn == 0
    if zero_based:
        i, j = i + 1, j + 1
    if is_res:
        v = 1.0 / v
    return i, j, v, zero_based


def network_graph(i, j, v):
    """network/pairwise.jl:31-50."""
    m = int(max(i.max(), j.max()))
    A = sp.coo_matrix((v, (i - 1, j - 1)), shape=(m, m)).tocsr()
    A = (A + A.T).tocsr()
    cc = connected_components(A)
    return laplacian(A), cc


def network_pairwise(cfg, inputs, solver="direct"):
    """network/pairwise.jl:4-29."""
    flags = cfg_flags(cfg)
    i, j, v, _ = load_graph(inputs["habitat_file"][1], cfg_bool(cfg, "habitat_map_is_resistances", "True"))
    fp = np.asarray(inputs["point_file"][1]).ravel().astype(np.int64)     # io.jl:74-82
    if fp.min() == 0:
        fp = fp + 1
    G, cc = network_graph(i, j, v)
    prob = GraphProblem(G, cc, fp, fp, set(), None, None, None, False, (i, j))
    return single_ground_all_pairs(prob, flags, solver)


# ----------------------------------------------------------------------------
# advanced mode (raster/advanced.jl)
# ----------------------------------------------------------------------------
def resolve_conflicts(sources, grounds, policy):
    """raster/advanced.jl:119-149 (incl. the :rmvall quirk pinned by
    test/internal.jl:130-135)."""
    sources = np.array(sources, dtype=np.float64)
    grounds = np.array(grounds, dtype=np.float64)
    finite = np.where(grounds < np.inf, grounds, 0.0)
    if np.count_nonzero(finite) == 0:
        finite = np.array([NODATA])
    conflicts = (sources != 0) & (grounds != 0)
    if conflicts.any():
        if policy in ("rmvsrc", "rmvall"):
            sources[conflicts] = 0
        elif policy == "rmvgnd":
            grounds[conflicts] = 0
    infc = (grounds == np.inf) & (sources > 0)
    grounds[infc] = 0
    return sources, grounds, finite


def multiple_solver(A, sources, grounds, finitegrounds, solver="direct"):
    """raster/advanced.jl:274-305: diag += finite grounds; rows/cols of Inf
    grounds removed and pinned to 0 V; solve; re-insert zeros."""
    A = sp.csr_matrix(A, dtype=np.float64)
    n = A.shape[0]
    if finitegrounds[0] != NODATA:
        A = (A + sp.diags(finitegrounds)).tocsr()
    inf = grounds == np.inf
    keep = np.nonzero(~inf)[0]
    As = A[keep][:, keep].tocsc()
    b = np.asarray(sources, dtype=np.float64)[keep]
    if solver in ("direct",):
        x = spla.splu(As).solve(b)
    elif solver == "cholmod":
        x = solve_cholmod_like(As, b)
    elif solver == "cg+amg":
        from . import amg
        ml = amg.smoothed_aggregation(sp.csr_matrix(As))      # defaults, raster/advanced.jl:308
        x, _ = amg.pcg(sp.csr_matrix(As), b, ml, rtol=1e-6, itmax=100_000)
    else:
        raise ValueError(solver)
    res = np.linalg.norm(As @ x - b) / np.linalg.norm(b)
    assert res < 1e-4                                            # raster/advanced.jl:310
    v = np.zeros(n)
    v[keep] = x
    return v


@dataclass
class AdvancedResult:
    voltages: np.ndarray            # per node (global)
    voltmap: np.ndarray | None = None
    curmap: np.ndarray | None = None
    node_currents: np.ndarray | None = None
    branch: tuple | None = None


def advanced_kernel(G, cc, sources, grounds, finitegrounds, nodemap=None, polymap=None,
                    cellmap=None, solver="direct"):
    """raster/advanced.jl:151-271 (advanced scenario: check_node = -1)."""
    G = sp.csr_matrix(G)
    n = G.shape[0]
    is_raster = nodemap is not None
    voltages = np.zeros(n)
    outvolt = np.zeros(nodemap.shape) if is_raster else None
    outcurr = np.zeros(nodemap.shape) if is_raster else None
    for c in cc:
        idx = np.asarray(c) - 1
        s_local, g_local = sources[idx].copy(), grounds[idx].copy()
        if s_local.sum() == 0 or g_local.sum() == 0:
            continue
        f_local = finitegrounds[idx] if finitegrounds[0] != NODATA else finitegrounds
        a_local = G[idx][:, idx].tocsr()
        v = multiple_solver(a_local, s_local, g_local, f_local, solver)
        voltages[idx] += v
        if is_raster:
            local_nodemap = construct_local_node_map(nodemap, c, polymap)
            outvolt += scatter_to_raster(voltages[idx], local_nodemap)       # out.jl:438-443
            outcurr += scatter_to_raster(get_node_currents(a_local, voltages[idx], f_local), local_nodemap)
    res = AdvancedResult(voltages, outvolt, outcurr)
    if not is_raster:
        res.node_currents = get_node_currents(G, voltages, finitegrounds)      # raster/advanced.jl:231
        res.branch = get_branch_currents_abs(G, voltages)
    return res


def _sources_grounds_raster(source_map, ground_map, nodemap, n, policy):
    """raster/advanced.jl:81-117 (raster branch)."""
    sources = np.zeros(n)
    grounds = np.zeros(n)
    r, c = colmajor_nonzero(source_map != 0)
    for a, b in zip(r, c):
        v = nodemap[a, b]
        if v != 0:
            sources[v - 1] += source_map[a, b]
    r, c = colmajor_nonzero(ground_map != 0)
    for a, b in zip(r, c):
        v = nodemap[a, b]
        if v != 0:
            grounds[v - 1] += ground_map[a, b]
    return resolve_conflicts(sources, grounds, policy)


def read_source_and_ground_maps(cfg, inputs, meta):
    """io.jl:256-326."""
    nrows, ncols = int(meta[1]), int(meta[0])

    def txt_to_rc(raw):
        raw = np.asarray(raw, dtype=np.float64)
        xll, yll, cs = float(meta[2]), float(meta[3]), float(meta[4])
        rr = np.ceil(nrows - (raw[:, 2] - yll) / cs).astype(np.int64)
        cc_ = np.ceil((raw[:, 1] - xll) / cs).astype(np.int64)
        return raw[:, 0], rr, cc_

    gk = inputs["ground_file"]
    if gk[0] == "grid":
        ground_map = read_polymap(gk[1], nodata_as=-1)
    else:
        val, rr, cc_ = txt_to_rc(gk[1])
        ground_map = np.full((nrows, ncols), NODATA)
        ground_map[rr - 1, cc_ - 1] = val
    sk = inputs["source_file"]
    if sk[0] == "grid":
        source_map = read_polymap(sk[1])
        source_map[source_map == NODATA] = 0
    else:
        val, rr, cc_ = txt_to_rc(sk[1])
        source_map = np.zeros((nrows, ncols))
        source_map[rr - 1, cc_ - 1] = val
    nod = ground_map == NODATA
    if cfg_bool(cfg, "ground_file_is_resistances", "True"):
        with np.errstate(divide="ignore"):
            ground_map = 1.0 / ground_map
    ground_map[nod] = 0
    if cfg_bool(cfg, "use_unit_currents"):
        source_map[source_map != 0] = 1
    if cfg_bool(cfg, "use_direct_grounds"):
        ground_map[ground_map != 0] = np.inf
    return source_map, ground_map


def raster_advanced(cfg, inputs, solver="direct"):
    """raster/advanced.jl:17-71."""
    flags = cfg_flags(cfg)
    cellmap, polymap, meta, _ = load_raster_inputs(cfg, inputs)
    source_map, ground_map = read_source_and_ground_maps(cfg, inputs, meta)
    nodemap = construct_node_map(cellmap, polymap)
    G = laplacian(construct_graph(cellmap, nodemap, flags["avg_res"], flags["four_neighbors"]))
    cc = connected_components(G)
    s, g, f = _sources_grounds_raster(source_map, ground_map, nodemap, G.shape[0], flags["policy"])
    return advanced_kernel(G, cc, s, g, f, nodemap, polymap, cellmap, solver)


def network_advanced(cfg, inputs, solver="direct"):
    """network/advanced.jl:1-51 + raster/advanced.jl:106-117 (network branch)."""
    flags = cfg_flags(cfg)
    i, j, v, zero_based = load_graph(inputs["habitat_file"][1], cfg_bool(cfg, "habitat_map_is_resistances", "True"))
    G, cc = network_graph(i, j, v)
    n = G.shape[0]

    def strengths(raw):                                           # io.jl:84-89
        raw = np.asarray(raw, dtype=np.float64).reshape(-1, 2).copy()
        if raw[:, 0].min() == 0 or zero_based:
            raw[:, 0] += 1
        return raw

    src = strengths(inputs["source_file"][1])
    gnd = strengths(inputs["ground_file"][1])
    if flags["grnd_file_is_res"]:
        with np.errstate(divide="ignore"):
            gnd[:, 1] = 1.0 / gnd[:, 1]
    sources = np.zeros(n); grounds = np.zeros(n)
    sources[src[:, 0].astype(np.int64) - 1] = src[:, 1]
    grounds[gnd[:, 0].astype(np.int64) - 1] = gnd[:, 1]
    s, g, f = resolve_conflicts(sources, grounds, flags["policy"])
    return advanced_kernel(G, cc, s, g, f, None, None, None, solver)


# ----------------------------------------------------------------------------
# one-to-all / all-to-one (raster/onetoall.jl) -- callers of the advanced kernel
# ----------------------------------------------------------------------------
def create_new_polymap_pointmap(gmap, polymap, points_rc, point_map):
    """raster/pairwise.jl:374-403 (the branch with a non-empty point_map)."""
    if polymap is None or np.size(polymap) == 0:
        return point_map.copy()
    newpoly = polymap.copy()
    no_polys = len(points_rc[2]) == len(np.unique(points_rc[2]))
    r, c = colmajor_nonzero(point_map != 0)
    if no_polys:
        k = polymap.max()
        for a, b in zip(r, c):
            if polymap[a, b] == 0:
                newpoly[a, b] = point_map[a, b] + k
    else:
        k = max(polymap.max(), point_map.max())
        for a, b in zip(r, c):
            v1, v2 = point_map[a, b], newpoly[a, b]
            if v2 == 0:
                newpoly[a, b] = k + v1
            elif v1 != v2:
                newpoly[newpoly == v2] = v1
    return newpoly


@dataclass
class OneToAllResult:
    resistances: np.ndarray                    # (P, 2): id, value  (raster/onetoall.jl:166)
    curmaps: dict = field(default_factory=dict)
    voltmaps: dict = field(default_factory=dict)
    cum_curmap: np.ndarray | None = None
    max_curmap: np.ndarray | None = None


def raster_one_to_all(cfg, inputs, solver="direct"):
    """raster/onetoall.jl:1-167; scenario one-to-all or all-to-one from cfg."""
    flags = cfg_flags(cfg)
    one_to_all = cfg.get("scenario") in ("one-to-all", "one_to_all")
    cellmap, polymap, meta, inc = load_raster_inputs(cfg, inputs)
    pk = inputs["point_file"]
    points_rc = read_point_map(pk[0], pk[1], meta)
    strengths = None
    if cfg_bool(cfg, "use_variable_source_strengths"):
        strengths = np.asarray(inputs["variable_source_file"][1], dtype=np.float64).reshape(-1, 2).copy()   # io.jl:84-89
        if strengths[:, 0].min() == 0:
            strengths[:, 0] += 1
    use_inc = inc is not None
    mode = 0 if (use_inc and inc.mode == "include") else 1
    if use_inc:
        keep = np.isin(points_rc[2], inc.point_ids)                   # prune_points!
        points_rc = tuple(a[keep] for a in points_rc)
        if strengths is not None:                                     # prune_strengths
            strengths = strengths[np.isin(strengths[:, 0], inc.point_ids)]
    rr, cc_, ids = points_rc
    point_map = np.zeros(cellmap.shape, dtype=np.int64)
    point_map[rr - 1, cc_ - 1] = ids
    points_unique = list(dict.fromkeys(int(p) for p in ids))
    newpoly = create_new_polymap_pointmap(cellmap, polymap, points_rc, point_map)
    nodemap = construct_node_map(cellmap, newpoly)
    a = construct_graph(cellmap, nodemap, flags["avg_res"], flags["four_neighbors"])
    cc = connected_components(a)
    G = laplacian(a)
    n_nodes = G.shape[0]
    unique_point_map = np.zeros(cellmap.shape, dtype=np.int64)
    for p in points_unique:
        ind = int(np.nonzero(ids == p)[0][0])
        unique_point_map[rr[ind] - 1, cc_[ind] - 1] = ids[ind]
    res = np.zeros(len(points_unique))
    out = OneToAllResult(resistances=None)
    out.cum_curmap = np.zeros(cellmap.shape)
    out.max_curmap = np.full(cellmap.shape, NODATA) if flags["write_max_cur_maps"] else None
    strength_map = np.zeros(cellmap.shape) if strengths is not None else None
    for i, n in enumerate(points_unique):
        pm = point_map.copy()
        nm, npoly = nodemap, newpoly
        strv = strengths[i, 1] if strengths is not None else 1.0
        if use_inc:
            for j in range(len(inc.point_ids)):
                if i != j and inc.mat[i, j] == mode:
                    pm[pm == int(inc.point_ids[j])] = 0
            npoly = create_new_polymap_pointmap(cellmap, polymap, points_rc, pm)
            nm = construct_node_map(cellmap, polymap)                 # (sic) raster/onetoall.jl:88
        if strengths is not None:
            tmp = pm[rr - 1, cc_ - 1]
            st2 = strengths.copy()
            st2[tmp == 0, 1] = 1
            strength_map[rr - 1, cc_ - 1] = st2[:, 1]
        if pm.sum() == n:
            res[i] = -1
            continue
        if one_to_all:
            source_map = np.where(unique_point_map == n, float(strv), 0.0)
            ground_map = np.where(pm == n, 0.0, pm.astype(np.float64))
            ground_map = np.where(ground_map > 0, np.inf, ground_map)
            policy = "rmvgnd"
        else:
            if strengths is not None:
                source_map = np.where(unique_point_map == n, 0.0, strength_map)
            else:
                source_map = np.where(unique_point_map != 0, 1.0, 0.0)
                source_map = np.where(pm == n, 0.0, source_map)
            ground_map = np.where(pm == n, np.inf, 0.0)
            policy = "rmvsrc"
        check_node = nm[rr[i] - 1, cc_[i] - 1]                        # (sic) indexes points_rc by i
        s_, g_, f_ = _sources_grounds_raster(source_map, ground_map, nm, n_nodes, policy)
        # advanced_kernel restricted to the component of check_node (raster/advanced.jl:186-188)
        volt = np.zeros(cellmap.shape)
        outvolt = np.zeros(cellmap.shape)
        outcurr = np.zeros(cellmap.shape)
        called = False
        Gc = sp.csr_matrix(G)
        for comp in cc:
            if check_node not in comp:
                continue
            idx = np.asarray(comp) - 1
            sl, gl = s_[idx].copy(), g_[idx].copy()
            if sl.sum() == 0 or gl.sum() == 0:
                continue
            fl = f_[idx] if f_[0] != NODATA else f_
            a_local = Gc[idx][:, idx].tocsr()
            v = multiple_solver(a_local, sl, gl, fl, solver)
            local_nodemap = construct_local_node_map(nm, comp, npoly)
            called = True
            outvolt += scatter_to_raster(v, local_nodemap)
            outcurr += scatter_to_raster(get_node_currents(a_local, v, fl), local_nodemap)
            lm = local_nodemap
            volt[lm != 0] = v[lm[lm != 0] - 1]
        if not called:
            res[i] = -1
        elif one_to_all:
            val = volt[source_map != 0] / source_map[source_map != 0]
            res[i] = -1 if np.isclose(val[0], 0) else val[0]
        else:
            res[i] = 0
        if flags["write_volt_maps"]:
            out.voltmaps[n] = outvolt
        if flags["write_cur_maps"] or flags["write_cum_cur_map_only"]:
            out.curmaps[n] = outcurr
        out.cum_curmap += outcurr                                      # raster/onetoall.jl:153-158
        if out.max_curmap is not None:
            out.max_curmap = np.maximum(out.max_curmap, outcurr)
    out.resistances = np.column_stack([points_unique, res])
    out.cum_curmap = np.where(out.cum_curmap < NODATA, NODATA, out.cum_curmap)
    return out


# ----------------------------------------------------------------------------
# fixtures helper
# ----------------------------------------------------------------------------
def load_case(npz, name):
    """(cfg dict, inputs dict key -> (kind, raw, meta), expected dict suffix -> array)."""
    cfg = json.loads(str(npz[f"{name}|cfg"]))
    inputs, expected = {}, {}
    for k in npz.files:
        parts = k.split("|")
        if parts[0] != name:
            continue
        if parts[1] == "in" and len(parts) == 3:
            inputs[parts[2]] = (str(npz[k + "|kind"]), npz[k], npz[k + "|meta"])
        elif parts[1] == "out":
            expected[parts[2]] = npz[k]
    return cfg, inputs, expected
