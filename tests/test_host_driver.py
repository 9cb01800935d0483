"""Host-side drivers (circuitscape_b200/core.py) against the reference goldens with
a CPU test double standing in for the device factor -- runs without a GPU.  The
same cases run through the real CUDA library in tests/test_gpu_parity.py."""
import numpy as np
import pytest

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
