"""Mid-size parity: a 128^3 two-phase sphere, linear-elastic CG, against the live CPU restatements — the NumPy oracle and the
independent multithreaded C++ port (oracle/cpu/fans_cpu.cpp).  At this size the marches are x-chunked, the persistent grids run
several waves and every FFT axis takes the multi-stage path, none of which the 16^3..32^3 cases reach.  Gates: BASELINE.md section 4."""
import numpy as np
import pytest

import fans_cpu
import fans_oracle as fo
import util
from util import ELASTIC, EP, rel_err
from fans_b200 import simple

pytestmark = pytest.mark.gpu
G0 = [0.001, -0.002, 0.003, 0.0015, -0.0025, 0.001]


def test_sphere128_linear_elastic_cg():
    n = 128
    ms = simple.sphere_microstructure(n)
    ctx = simple.linear_elastic_context(ms, [1.0, 1.0, 1.0], [62.5, 222.222], [28.8462, 166.6667], "HEX8")
    ctx.set_gradient(G0)
    res = ctx.solve("cg", 100, 1e-10, "Linfinity", "absolute")
    sg, ug = ctx.homogenized_stress(), ctx.download("u")
    ctx.close()
    cpu = fans_cpu.two_phase_elastic(ms, [1.0, 1.0, 1.0], [62.5, 222.222], [28.8462, 166.6667])
    rc = cpu.solve(G0, 100, 1e-10)
    assert abs(res["iters"] - rc["iters"]) <= 1, (res["iters"], rc["iters"])
    assert rel_err(sg, rc["sigma"]) < 1e-9
    assert rel_err(ug, cpu.u()) < 1e-8
    k = min(res["iters"], rc["iters"]) + 1
    assert rel_err(res["err_all"][:k], rc["err_all"][:k]) < 1e-6
    cpu.close()
    sol = fo.OracleSolver(ms, [1.0, 1.0, 1.0], "mechanical", ELASTIC, "HEX8", "cg", "small", EP, 100)
    sol.set_gradient(G0)
    sol.solve()
    assert abs(res["iters"] - sol.iter) <= 1, (res["iters"], sol.iter)
    assert rel_err(sg, sol.get_homogenized_stress()) < 1e-9
    assert rel_err(ug, sol.u) < 1e-8


def test_anisotropic_grid_256x64x128_vs_cpu_port():
    """long x marches (xchunk < n0), different length per axis, non-cubic voxels"""
    shape = (256, 64, 128)
    ms = util.two_phase_ms(0, 21, shape)
    L = [2.0, 1.0, 1.5]
    ctx = simple.linear_elastic_context(ms, L, [62.5, 222.222], [28.8462, 166.6667], "HEX8")
    ctx.set_gradient(G0)
    res = ctx.solve("cg", 200, 1e-10, "Linfinity", "absolute")
    sg, ug = ctx.homogenized_stress(), ctx.download("u")
    ctx.close()
    cpu = fans_cpu.two_phase_elastic(ms, L, [62.5, 222.222], [28.8462, 166.6667])
    rc = cpu.solve(G0, 200, 1e-10)
    assert abs(res["iters"] - rc["iters"]) <= 1
    assert rel_err(sg, rc["sigma"]) < 1e-9
    assert rel_err(ug, cpu.u()) < 1e-8
    cpu.close()
