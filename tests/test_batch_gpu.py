"""Batched linear solves (fans_solve_batch): the n_str load cases of Solver::get_homogenized_tangent
(/root/reference/include/solver.h:739-778) as lanes of ONE CG loop.  Every lane must reproduce the single solve with the same
macroscopic strain started from u = 0 — same iteration count, error history, displacement and homogenized stress (the arithmetic per
lane is the single solve's; only the order of the deterministic partial sums of the stencil's <d, K d> can differ, hence 1e-12 and
not bit equality) — and the single solve is what the oracle tests pin to the reference."""
import numpy as np
import pytest

import fans_oracle as fo
import util
from util import ELASTIC, EP, rel_err
from fans_b200 import simple
from fans_b200._lib import FansError

pytestmark = pytest.mark.gpu
BULK, SHEAR = [62.5, 222.222], [28.8462, 166.6667]


def _single(ctx, g0, n_it, tol, measure="Linfinity", err_type="absolute"):
    ctx.zero("u")
    ctx.set_gradient(g0)
    res = ctx.solve("cg", n_it, tol, measure, err_type)
    return res, ctx.homogenized_stress(), ctx.download("u")


def _check_lanes(ctx, macro, n_it, tol, measure="Linfinity", err_type="absolute", rtol=1e-12):
    res, sig = ctx.solve_batch(macro, n_it, tol, measure, err_type)
    us = []
    for l in range(len(macro)):
        ctx.batch_displacement(l, "u_prev")
        us.append(ctx.download("u_prev"))
    for l, g0 in enumerate(macro):
        r1, s1, u1 = _single(ctx, g0, n_it, tol, measure, err_type)
        assert res[l]["iters"] == r1["iters"], (l, res[l]["iters"], r1["iters"])
        if r1["iters"] == 0:
            assert np.abs(us[l]).max() == 0.0
            continue
        assert rel_err(res[l]["err_all"], r1["err_all"]) < 1e-9, l   # late entries sit 10 orders below the first: rounding of the sums
        assert rel_err(sig[l], s1) < rtol, (l, sig[l], s1)
        assert rel_err(us[l], u1) < 1e-11, l
    return res, sig


def test_batch_six_unit_load_cases_32():
    """config 2 of BASELINE.json in small: six unit strains, relative L-infinity error"""
    ms = simple.sphere_microstructure(32)
    ctx = simple.linear_elastic_context(ms, [1.0, 1.0, 1.0], BULK, SHEAR, "HEX8")
    res, sig = _check_lanes(ctx, np.eye(6), 100, 1e-8, "Linfinity", "relative")
    # the effective stiffness the six lanes give is symmetric and lies between the phases'
    Cbar = sig.T
    assert np.abs(Cbar - Cbar.T).max() < 1e-6 * np.abs(Cbar).max()
    assert 28.8462 * 2 < Cbar[3, 3] < 166.6667 * 2
    ctx.close()


def test_batch_against_oracle_and_state_untouched():
    """lane results against the NumPy oracle; the context's own u / gradient survive a batched solve"""
    ms = util.two_phase_ms(3, 17, (16, 16, 16))
    ctx = simple.linear_elastic_context(ms, [1.0, 1.0, 1.0], BULK, SHEAR, "HEX8")
    g_own = [0.002, 0.0, -0.001, 0.0005, 0.0, 0.001]
    ctx.set_gradient(g_own)
    ctx.solve("cg", 100, 1e-10, "Linfinity", "absolute")
    u_own = ctx.download("u")
    macro = np.array([[0.001, -0.002, 0.003, 0.0015, -0.0025, 0.001], [0.0, 0.0, 0.0, 0.004, 0.0, 0.0]])
    res, sig = ctx.solve_batch(macro, 100, 1e-10, "Linfinity", "absolute")
    assert np.array_equal(ctx.download("u"), u_own)
    assert np.allclose(ctx.get_gradient(), g_own, rtol=0, atol=0)
    for l in range(2):
        sol = fo.OracleSolver(ms, [1.0, 1.0, 1.0], "mechanical", ELASTIC, "HEX8", "cg", "small", EP, 100)
        sol.set_gradient(list(macro[l]))
        sol.solve()
        assert abs(res[l]["iters"] - sol.iter) <= 1
        assert rel_err(sig[l], sol.get_homogenized_stress()) < 1e-9
        ctx.batch_displacement(l, "u_prev")
        assert rel_err(ctx.download("u_prev"), sol.u) < 1e-8
    ctx.close()


def test_batch_lanes_converge_at_different_iterations():
    """absolute tolerance + strains of very different size: lanes freeze one by one (one of them before the first iteration)
    and keep exactly what the single solve leaves"""
    ms = util.two_phase_ms(5, 11, (32, 16, 64))
    ctx = simple.linear_elastic_context(ms, [2.0, 1.0, 1.5], BULK, SHEAR, "HEX8")
    base = np.array([0.001, -0.002, 0.003, 0.0015, -0.0025, 0.001])
    macro = np.stack([base, 1e-4 * base, np.zeros(6), 1e-8 * base[::-1], 10.0 * base])
    res, _ = _check_lanes(ctx, macro, 200, 1e-9, "L2", "absolute")
    its = [r["iters"] for r in res]
    assert its[2] == 0 and len(set(its)) >= 3, its
    ctx.close()


def test_batch_thermal_and_bbar_and_polycrystal():
    # scalar problem (h = 1), three unit gradients
    ms = simple.sphere_microstructure(32)
    ctx = simple.linear_thermal_context(ms, [1.0, 1.0, 1.0], [1.0, 10.0], "HEX8")
    _check_lanes(ctx, np.eye(3), 100, 1e-10, "Linfinity", "absolute")
    ctx.close()
    # BBAR elements (non-isotropic stencil pattern)
    ms = util.two_phase_ms(1, 9, (16, 32, 16))
    ctx = simple.linear_elastic_context(ms, [1.0, 2.0, 1.0], BULK, SHEAR, "BBAR")
    _check_lanes(ctx, np.eye(6)[:3] * 0.01, 100, 1e-10, "Linfinity", "absolute")
    ctx.close()
    # 12 grains with their own rotated cubic tensor: the coefficient-table path of the stencil
    labels = simple.voronoi_labels((32, 32, 32), 12, seed=5)
    ctx = simple.linear_elastic_tensor_context(labels, [1.0, 1.0, 1.0], simple.rotated_cubic_tangents(12), "HEX8")
    _check_lanes(ctx, np.eye(6) * 0.01, 200, 1e-9, "Linfinity", "absolute")
    ctx.close()


def test_batch_midsize_chunked_marches():
    """128 x 64 x 128: the stencil marches are x-chunked and the update runs several waves per lane"""
    ms = util.two_phase_ms(0, 21, (128, 64, 128))
    ctx = simple.linear_elastic_context(ms, [2.0, 1.0, 1.5], BULK, SHEAR, "HEX8")
    _check_lanes(ctx, np.eye(6)[[0, 3, 5]] * 0.01, 200, 1e-10, "Linfinity", "absolute")
    ctx.batch_release()
    ctx.close()


def test_batch_refused_where_the_reference_loop_remains():
    ms = simple.fiber_microstructure(16, n_fibers=4)
    ctx = simple.j2_fiber_context(ms, [1.0, 1.0, 1.0])
    with pytest.raises(FansError, match="linear material"):
        ctx.solve_batch(np.eye(6) * 1e-3, 10, 1e-8)
    ctx.close()
    ms = util.two_phase_ms(0, 3, (12, 16, 16))   # not a power of two
    ctx = simple.linear_elastic_context(ms, [1.0, 1.0, 1.0], BULK, SHEAR, "HEX8")
    with pytest.raises(FansError, match="power-of-two"):
        ctx.solve_batch(np.eye(6) * 1e-3, 10, 1e-8)
    ctx.close()
