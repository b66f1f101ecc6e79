"""Grids whose dimensions are not powers of two (the reference's FFTW plans take any n_x, n_y, n_z; src/reader.cpp:300-305): the
Bluestein passes of fans_b200/csrc/fft_any.cu, the element sweeps / stencil on partial tiles, odd sizes included, against the oracle."""
import numpy as np
import pytest

import fans_oracle as fo
import util
from util import ELASTIC, THERMAL, EP, rel_err

pytestmark = pytest.mark.gpu
G0 = {3: [0.01, 0.02, -0.01], 6: [0.001, -0.002, 0.003, 0.0015, -0.0025, 0.001]}
SHAPES = [(12, 20, 24), (9, 15, 10), (10, 12, 7), (24, 6, 100), (5, 7, 10), (6, 7, 9)]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("problem,materials", [("thermal", THERMAL), ("mechanical", ELASTIC)])
def test_operators_and_cg_on_any_grid(shape, problem, materials):
    ms = util.two_phase_ms(0, 31, shape)
    sol = fo.OracleSolver(ms, [1.0, 1.5, 2.0], problem, materials, "HEX8", "cg", "small", EP, 200)
    ctx = util.ctx_from_oracle(sol)
    rng = np.random.default_rng(2)
    r = rng.standard_normal(ctx.field_shape)
    ctx.upload("r", r)
    ctx.convolution("r", "s")
    assert rel_err(ctx.download("s"), sol.convolution(r)) < 1e-10
    ctx.convolution("r", "r")   # in place (SolverFP, include/solverFP.h:29)
    assert rel_err(ctx.download("r"), sol.convolution(r)) < 1e-10
    u = rng.standard_normal(ctx.field_shape) * 1e-3
    ctx.upload("u", u)
    ctx.apply_linear("rnew", "u")
    assert rel_err(ctx.download("rnew"), sol.apply_linear(u)) < 1e-11
    g0 = np.array(G0[sol.n_str])
    sol.set_gradient(g0)
    ctx.set_gradient(g0)
    ctx.residual("r", "u")
    assert rel_err(ctx.download("r"), sol.compute_residual(u)) < 1e-11
    ref = float((u * sol.compute_residual(u)).sum())
    assert abs(ctx.dot("u", "r") - ref) < 1e-11 * abs(ref) + 1e-24      # odd element counts: the vector passes cover the last value
    assert abs(ctx.norm("u", "L1") - np.abs(u).sum()) < 1e-12 * np.abs(u).sum()
    ctx.zero("u")
    sol.solve()
    res = ctx.solve("cg", 200, EP["tolerance"], EP["measure"], EP["type"])
    assert abs(res["iters"] - sol.iter) <= 1, (res["iters"], sol.iter)
    assert rel_err(ctx.homogenized_stress(), sol.get_homogenized_stress()) < 1e-9
    assert rel_err(ctx.download("u"), sol.u) < 1e-8
    ctx.close()


def test_j2_plasticity_on_a_100_like_grid():
    """nonlinear path (history, line search) on a non-2^k grid: 20 x 12 x 18"""
    import test_nonlinear_gpu as tn
    ms = util.two_phase_ms(0, 17, (20, 12, 18))
    tn.compare(tn.cfg_for([tn.MODELS["j2_lin"], tn.EL1], "HEX8", "cg", tn.LOAD[:2]), ms, ("plastic_strain",))


def test_odd_frequency_count_is_refused():
    """n_x n_y (n_z/2+1) odd: the reference's pairwise Gamma loop drops the last frequency (solver.h:391-400) — refused loudly"""
    from fans_b200 import _lib as L
    with pytest.raises(L.FansError, match="skips the last frequency"):
        L.Context((5, 7, 9), [1, 1, 1], 1, 3, "HEX8")


def test_slabs_need_power_of_two():
    from fans_b200 import _lib as L

    class FakeComm:
        world_size, rank, handle = 2, 0, None
    with pytest.raises(L.FansError, match="power-of-two"):
        L.Context((12, 12, 12), [1, 1, 1], 3, 6, "HEX8", comm=FakeComm())
