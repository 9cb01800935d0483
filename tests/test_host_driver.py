"""Host-side drivers (circuitscape_b200/core.py) against the reference goldens with
a CPU test double standing in for the device factor -- runs without a GPU.  The
same cases run through the real CUDA library in tests/test_gpu_parity.py."""
import numpy as np
import pytest
import scipy.sparse as sp

import circuitscape_b200 as cb
from circuitscape_b200 import solver as S

from oracle import circuitscape_oracle as co

from . import cases
from .fake_factor import FakeFactor


@pytest.fixture(autouse=True)
def fake_device(monkeypatch):
    monkeypatch.setattr(S, "construct_cholesky_factor", lambda m, s, **kw: FakeFactor(m, s, **kw))
    monkeypatch.setattr(S, "multiple_solve",
                        lambda s, m, b: FakeFactor(m, s).solve_rhs(np.asarray(b))[0])


@pytest.mark.parametrize("i", range(1, 18))
def test_raster_pairwise_driver(golden, i):
    r, exp = cases.run_raster_pairwise(golden, f"sgVerify{i}", cb.CUDASolver())
    cases.check_raster_pairwise(r, exp, rel=1e-7)


@pytest.mark.parametrize("i", range(1, 4))
def test_network_pairwise_driver(golden, i):
    prob, flags, exp = cases.network_pairwise_problem(golden, f"sgNetworkVerify{i}", cb.CUDASolver())
    cases.check_network_pairwise(cb.single_ground_all_pairs(prob, flags), exp)


@pytest.mark.parametrize("name", [f"mgVerify{i}" for i in range(1, 7)] +
                         [f"mgNetworkVerify{i}" for i in range(1, 4)])
def test_advanced_driver(golden, name):
    prob, flags, exp = cases.advanced_problem(golden, name, cb.CUDASolver())
    cases.check_advanced(cb.advanced_kernel(prob, flags), exp, flags)


ONE_TO_ALL = [f"oneToAllVerify{i}" for i in range(1, 14)] + [f"allToOneVerify{i}" for i in range(1, 13)]


@pytest.mark.parametrize("name", ONE_TO_ALL)
def test_onetoall_driver(golden, name):
    """src/raster/onetoall.jl through the advanced kernel (test/test_utils.jl:123-139)."""
    data, flags, cfg, exp = cases.onetoall_problem(golden, name)
    fl = co.cfg_flags(cfg)
    r = cb.onetoall_kernel(data, flags, cfg, solver=cb.CUDASolver(),
                           four_neighbors=fl["four_neighbors"], avg_res=fl["avg_res"])
    cases.check_onetoall(r, exp, flags)


@pytest.mark.parametrize("name", [n for n in ONE_TO_ALL if n.startswith("allToOne")])
def test_alltoone_driver_batched(golden, name):
    """CUDASolver(batch_all_to_one=True): iterations of a component as columns of one sparse-RHS
    batch on the singular operator; same goldens."""
    data, flags, cfg, exp = cases.onetoall_problem(golden, name)
    fl = co.cfg_flags(cfg)
    r = cb.onetoall_kernel(data, flags, cfg, solver=cb.CUDASolver(batch_all_to_one=True),
                           four_neighbors=fl["four_neighbors"], avg_res=fl["avg_res"])
    cases.check_onetoall(r, exp, flags)


@pytest.mark.parametrize("name", [n for n in ONE_TO_ALL if n.startswith("oneToAll")])
def test_onetoall_driver_batched(golden, name):
    """CUDASolver(batch_one_to_all=True): the iterations of a component as columns of one batch on the
    Laplacian with every focal row/column removed (block elimination of the live focal node)."""
    data, flags, cfg, exp = cases.onetoall_problem(golden, name)
    fl = co.cfg_flags(cfg)
    r = cb.onetoall_kernel(data, flags, cfg, solver=cb.CUDASolver(batch_one_to_all=True),
                           four_neighbors=fl["four_neighbors"], avg_res=fl["avg_res"])
    cases.check_onetoall(r, exp, flags)


def test_batching_is_transparent(golden):
    """cholmod_batch_size only changes how pairs are grouped (src/core.jl:448-452)."""
    a, exp = cases.run_raster_pairwise(golden, "sgVerify4", cb.CUDASolver(bs=1000))
    b, _ = cases.run_raster_pairwise(golden, "sgVerify4", cb.CUDASolver(bs=4))
    assert np.allclose(a.resistances, b.resistances, atol=1e-12)
    assert np.allclose(a.cum_curmap, b.cum_curmap, atol=1e-12)


def test_solver_selection():
    assert isinstance(cb.get_solver({"solver": "cuda", "cholmod_batch_size": "16"}), cb.CUDASolver)
    assert cb.get_solver({"solver": "b200", "precision": "single"}).dtype == np.float32
    with pytest.raises(ValueError):
        cb.get_solver({"solver": "cholmod"})


# ---------------------------------------------------------------------------
# network-mode synthetic graph (config C5) and the batched all-to-one identity
# ---------------------------------------------------------------------------
def test_power_law_laplacian_shape():
    import scipy.sparse as sp
    from circuitscape_b200 import graph
    L = graph.power_law_laplacian(20000, m=5, seed=11)
    assert L.shape == (20000, 20000)
    assert abs(L - L.T).max() == 0
    assert np.abs(np.asarray(L.sum(axis=1))).max() < 1e-9
    deg = np.diff(L.indptr) - 1
    assert deg.min() >= 1 and deg.max() > 30 * deg.mean()          # hubs
    assert len(graph.connected_components(L)) == 1
    off = L - sp.diags(L.diagonal())
    assert off.data.max() < 0 and -off.data.min() <= 1.0 and -off.data.max() >= 0.1


class _PinvFactor:
    """singular-Laplacian solve by dense pseudo-inverse (what CG returns up to a constant)"""

    def __init__(self, L):
        self.n = L.shape[0]
        self.io_dtype = np.float64
        self.P = np.linalg.pinv(L.toarray())

    def solve_rhs(self, rhs, rtol=None):
        k = rhs.shape[1]
        return self.P @ rhs + 3.25, np.zeros(k, dtype=np.int64), np.zeros(k)


def test_all_to_one_batched_equals_grounded_solves():
    """one singular operator + shift == deleting the ground's row/column per iteration
    (src/raster/advanced.jl:286-300 with the sources of src/raster/onetoall.jl:110-118)"""
    from circuitscape_b200 import graph, core
    L = graph.power_law_laplacian(400, m=3, seed=5)
    focal = graph.focal_nodes(400, 6, seed=2)
    V, it, rr, cols = core.all_to_one_batched(_PinvFactor(L), focal)
    assert V.shape == (400, 6) and list(cols) == list(range(6))
    for c, f in enumerate(focal):
        keep = np.setdiff1d(np.arange(400), [f])
        b = np.zeros(400); b[focal] = 1.0
        v = np.zeros(400)
        v[keep] = np.linalg.solve(L[keep][:, keep].toarray(), b[keep])
        assert np.abs(V[:, c] - v).max() < 1e-9 * np.abs(v).max()
    V2, _, _, cols2 = core.all_to_one_batched(_PinvFactor(L), focal, shard=(1, 4))
    assert list(cols2) == [1, 5] and np.allclose(V2, V[:, [1, 5]])


# ---------------------------------------------------------------------------
# compute_omniscape_current (src/utils.jl:145-257)
# ---------------------------------------------------------------------------
def test_compute_omniscape_current_reference_example():
    """the reference's own call (test/internal.jl:5-43) plus a check against the oracle's
    advanced-mode raster path on the same window."""
    conductance = np.array([[1, 5, 1.], [2, 1, 1], [9, 1, 6]])
    source = np.array([[1, 0, 0.], [0, 0, 0], [0, 1, 0]])
    ground = np.array([[0, 0, 1.], [0, 0, 0], [0, 0, 0]])
    cs_cfg = {"ground_file_is_resistances": "True", "use_direct_grounds": "False", "output_file": "temp",
              "write_cum_cur_map_only": "False", "scenario": "Advanced", "suppress_messages": "True",
              "connect_four_neighbors_only": "False", "solver": "cuda", "cholmod_batch_size": "1000",
              "data_type": "raster"}
    cur = cb.compute_omniscape_current(conductance, source, ground, cs_cfg)
    assert cur.shape == (3, 3) and np.all(np.isfinite(cur)) and cur.min() >= 0
    # 2 A injected, all of it leaves through the single ground cell (conductance 1 to earth)
    assert abs(cur[0, 2] - 2.0) < 1e-9
    # oracle: same graph, same sources/grounds
    nodemap = co.construct_node_map(conductance, None)
    G = co.laplacian(co.construct_graph(conductance, nodemap, False, False))
    s, g, f = co._sources_grounds_raster(source, ground, nodemap, G.shape[0], "rmvsrc")
    M = (G + sp.diags(f)).tocsc()
    import scipy.sparse.linalg as spla
    v = spla.splu(M).solve(s)
    ref = np.zeros((3, 3))
    nc = co.get_node_currents(G, v, f)
    ref[nodemap > 0] = nc[nodemap[nodemap > 0] - 1]
    assert np.abs(cur - ref).max() < 1e-9


def test_compute_omniscape_current_window_with_holes():
    rng = np.random.default_rng(5)
    g = rng.uniform(0.5, 2.0, (12, 9)); g[rng.random(g.shape) < 0.15] = -9999.0
    src = np.zeros_like(g); src[2, 3] = 1.0; src[9, 7] = 0.5
    gnd = np.zeros_like(g); gnd[6, 1] = 2.0; gnd[11, 8] = 1.0
    for k in (src, gnd):
        k[g <= 0] = 0
    cur = cb.compute_omniscape_current(g, src, gnd, {"connect_four_neighbors_only": "True", "solver": "cuda"})
    assert cur.shape == g.shape and np.all(cur[g <= 0] == 0) and cur.max() > 0


@pytest.mark.parametrize("i", [1, 2, 5, 9, 16])
def test_raster_pairwise_driver_superposed(golden, i):
    """CUDASolver(superpose=True): the driver hands each component's pairs to
    solve_pairs_superposed as (nodes, pi, pj); outputs must still meet the goldens."""
    r, exp = cases.run_raster_pairwise(golden, f"sgVerify{i}", cb.CUDASolver(superpose=True))
    cases.check_raster_pairwise(r, exp, rel=1e-7)


def test_network_pairwise_driver_superposed(golden):
    prob, flags, exp = cases.network_pairwise_problem(golden, "sgNetworkVerify2", cb.CUDASolver(superpose=True))
    cases.check_network_pairwise(cb.single_ground_all_pairs(prob, flags), exp)


# ---- round-2 advisor findings -----------------------------------------------------------------
def test_network_log_transform_keeps_raw_cumulative_currents(golden):
    """Only raster maps are log-transformed (src/out.jl:96); the network branch accumulates raw node
    currents (src/out.jl:48-88), with or without `log_transform_maps`."""
    prob, flags, exp = cases.network_pairwise_problem(golden, "sgNetworkVerify1", cb.CUDASolver())
    plain = cb.single_ground_all_pairs(prob, flags)
    prob2, flags2, _ = cases.network_pairwise_problem(golden, "sgNetworkVerify1", cb.CUDASolver())
    flags2.outputflags.log_transform_maps = True
    logged = cb.single_ground_all_pairs(prob2, flags2)
    assert np.allclose(plain.cum_node, logged.cum_node, rtol=0, atol=1e-12)
    cases.check_network_pairwise(logged, exp)


def test_network_cumulative_branch_currents_and_row_order(golden):
    """cum_branch follows `coords` (src/utils.jl:132-142) and matches the golden file; per-pair
    branch rows come column-major like `_convert_to_3col` (src/out.jl:128-148)."""
    prob, flags, exp = cases.network_pairwise_problem(golden, "sgNetworkVerify1", cb.CUDASolver())
    r = cb.single_ground_all_pairs(prob, flags)
    v = exp["branch_currents_cum.txt"].copy()
    v[:, :2] += 1
    mine = np.column_stack([prob.coords[0], prob.coords[1], r.cum_branch])
    mine = mine[~np.isclose(mine[:, 2], 0.0, atol=1e-6)]
    assert mine.shape == v.shape
    assert np.sum((cases.sorted_rows(mine) - cases.sorted_rows(v)) ** 2) < 1e-6
    gr, gc, _ = next(iter(r.branch.values()))
    assert np.all(gr < gc)
    key = gc.astype(np.int64) * (gc.max() + 1) + gr
    assert np.all(np.diff(key) > 0)


def test_unknown_branch_raises():
    from circuitscape_b200.core import _BranchIndex
    idx = _BranchIndex((np.array([1, 2]), np.array([2, 3])))
    cum = np.zeros(2)
    idx.add(cum, (np.array([2, 3]), np.array([1, 2]), np.array([0.5, 0.25])))      # reversed edges are found
    assert np.allclose(cum, [0.5, 0.25])
    with pytest.raises(KeyError):
        idx.add(cum, (np.array([1]), np.array([3]), np.array([1.0])))


class _Sink:
    def __init__(self):
        self.volt, self.cur, self.net = {}, {}, {}

    def voltmap(self, key, grid):
        self.volt[key] = grid.copy()

    def curmap(self, key, grid):
        self.cur[key] = grid.copy()

    def network(self, key, comp, v, cur, branch):
        self.net[key] = (comp, v, cur, branch)


def test_sink_streams_maps_instead_of_keeping_them(golden):
    """With a sink the per-pair maps are handed over as each batch finishes and not retained
    (the reference writes and drops them inside postprocess, src/core.jl:655-683)."""
    kept, exp = cases.run_raster_pairwise(golden, "sgVerify1", cb.CUDASolver())
    sink = _Sink()
    streamed, _ = cases.run_raster_pairwise(golden, "sgVerify1", cb.CUDASolver(), sink=sink)
    assert not streamed.curmaps and not streamed.voltmaps
    assert sink.cur.keys() == kept.curmaps.keys() and sink.volt.keys() == kept.voltmaps.keys()
    for k in kept.curmaps:
        assert np.array_equal(sink.cur[k], kept.curmaps[k])
    for k in kept.voltmaps:
        assert np.array_equal(sink.volt[k], kept.voltmaps[k])
    assert np.array_equal(streamed.cum_curmap, kept.cum_curmap)
    prob, flags, _ = cases.network_pairwise_problem(golden, "sgNetworkVerify2", cb.CUDASolver())
    nsink = _Sink()
    r = cb.single_ground_all_pairs(prob, flags, sink=nsink)
    assert not r.curmaps and not r.branch and nsink.net


@pytest.mark.parametrize("name", ONE_TO_ALL)
def test_onetoall_driver_resident_grounds(golden, name):
    """CUDASolver(resident_grounds=True): one factor per component kept across the iterations of the
    one-to-all / all-to-one loop, the grounds moved by set_grounds (identity rows instead of deleted
    rows, src/raster/advanced.jl:274-305); same goldens."""
    data, flags, cfg, exp = cases.onetoall_problem(golden, name)
    fl = co.cfg_flags(cfg)
    r = cb.onetoall_kernel(data, flags, cfg, solver=cb.CUDASolver(resident_grounds=True),
                           four_neighbors=fl["four_neighbors"], avg_res=fl["avg_res"])
    cases.check_onetoall(r, exp, flags)
