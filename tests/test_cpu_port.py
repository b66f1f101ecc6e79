"""The multithreaded C++ CPU restatement (oracle/cpu/fans_cpu.cpp: the timed CPU arm of bench.py) against the NumPy oracle, which is
itself pinned by the reference's known-answer tests (tests/test_oracle_kat.py).  Same algorithm, independent code: own FFT, own
element loops, threads over x-slabs."""
import numpy as np
import pytest

import fans_cpu
import fans_oracle as fo
import util
from util import rel_err

G0 = [0.001, -0.002, 0.003, 0.0015, -0.0025, 0.001]


def oracle_for(ms, L, n_it=200):
    return fo.OracleSolver(ms, L, "mechanical", util.ELASTIC, "HEX8", "cg", "small", util.EP, n_it)


@pytest.mark.parametrize("shape,threads", [((16, 16, 16), 1), ((8, 16, 32), 3), ((32, 16, 8), 5)])
def test_cpu_port_matches_oracle(shape, threads):
    ms = util.two_phase_ms(0, 3, shape)
    L = [1.0, 1.5, 2.0]
    sol = oracle_for(ms, L)
    sol.set_gradient(G0)
    sol.solve()
    cpu = fans_cpu.two_phase_elastic(ms, L, [62.5, 222.222], [28.8462, 166.6667], threads)
    res = cpu.solve(G0, 200, 1e-10)
    assert res["iters"] == sol.iter
    assert rel_err(res["err_all"], sol.err_all[: sol.iter + 1]) < 1e-6
    assert rel_err(res["sigma"], sol.get_homogenized_stress()) < 1e-10
    assert rel_err(cpu.u(), sol.u) < 1e-9
    cpu.close()


def test_cpu_port_operators():
    """convolution and K.d on random fields, thermal (h = 1) included through the generic tangent interface"""
    ms = util.two_phase_ms(0, 4, (16, 8, 16))
    sol = oracle_for(ms, [1.0, 1.0, 1.0], 0)
    cpu = fans_cpu.two_phase_elastic(ms, [1.0, 1.0, 1.0], [62.5, 222.222], [28.8462, 166.6667], 2)
    rng = np.random.default_rng(1)
    f = rng.standard_normal(ms.shape + (3,))
    assert rel_err(cpu.convolution(f), sol.convolution(f)) < 1e-11
    assert rel_err(cpu.apply_linear(f), sol.apply_linear(f)) < 1e-12
    cpu.close()
    k = np.array([1.0, 10.0])
    solt = fo.OracleSolver(ms, [1.0, 1.0, 1.0], "thermal", util.THERMAL, "HEX8", "cg", "small", util.EP, 100)
    solt.set_gradient([0.01, 0.02, -0.01])
    solt.solve()
    cput = fans_cpu.CpuSolver(ms, [1.0, 1.0, 1.0], np.stack([ki * np.eye(3) for ki in k]), np.eye(3) * k.mean(), 1, 2)
    r = cput.solve([0.01, 0.02, -0.01], 100, 1e-10)
    assert r["iters"] == solt.iter
    assert rel_err(r["sigma"], solt.get_homogenized_stress()) < 1e-10
    assert rel_err(cput.u(), solt.u) < 1e-9
    cput.close()
