"""The linear CG iteration replayed as a CUDA graph (small grids: launch-latency bound; solve.cu, fans_ctx::IterGraph).  A replay
launches the same kernels with the same arguments as the plain loop, so everything must be bit-identical to FANS_GRAPH=0 — across
solves that reuse the cached graphs, after the direction ping-pong ended on either buffer, and after the materials changed (the
stencil coefficients are kernel parameters frozen into the graph: it has to be rebuilt)."""
import os

import numpy as np
import pytest

import util
from fans_b200 import simple

pytestmark = pytest.mark.gpu
BULK, SHEAR = [62.5, 222.222], [28.8462, 166.6667]
LOADS = [[0.001, -0.002, 0.003, 0.0015, -0.0025, 0.001], [0.0, 0.004, 0.0, 0.0, 0.001, 0.0], [0.002, 0.0, 0.0, 0.0, 0.0, -0.003]]


def _run(graph, ms, loads, tol=1e-10, bulk2=None):
    os.environ["FANS_GRAPH"] = "1" if graph else "0"
    try:
        ctx = simple.linear_elastic_context(ms, [1.0, 1.0, 1.0], BULK, SHEAR, "HEX8")
        out = []
        for i, g in enumerate(loads):
            if bulk2 is not None and i == len(loads) - 1:   # new materials on the same context: coefficients change, pointers do not
                ctx.set_materials(simple.elastic_phase_descs(bulk2, SHEAR))
            ctx.set_gradient(g)
            l0 = ctx.launch_count()
            r = ctx.solve("cg", 200, tol, "Linfinity", "absolute")   # keeps u of the previous solve as the start, like the reference
            out.append((r["iters"], r["err_all"].copy(), ctx.homogenized_stress(), ctx.download("u"), ctx.launch_count() - l0, r["fft_ms"]))
        ctx.close()
        return out
    finally:
        os.environ.pop("FANS_GRAPH", None)


@pytest.mark.parametrize("shape", [(32, 32, 32), (16, 32, 64)])
def test_graph_replay_is_bit_identical(shape):
    ms = util.two_phase_ms(0, 5, shape)
    a, b = _run(True, ms, LOADS), _run(False, ms, LOADS)
    for (ia, ea, sa, ua, la, fa), (ib, eb, sb, ub, lb, fb) in zip(a, b):
        assert ia == ib and ia > 5
        assert np.array_equal(ea, eb)
        assert np.array_equal(sa, sb)
        assert np.array_equal(ua, ub)
        assert la == lb          # a replay counts the kernels it launches
        assert fa > 0.0          # the convolution time of the plain iterations stands in for the replayed ones


def test_graph_rebuilt_after_new_materials():
    ms = util.two_phase_ms(0, 6, (32, 32, 32))
    a = _run(True, ms, LOADS, bulk2=[80.0, 150.0])
    b = _run(False, ms, LOADS, bulk2=[80.0, 150.0])
    for (ia, ea, sa, ua, _, _), (ib, eb, sb, ub, _, _) in zip(a, b):
        assert ia == ib
        assert np.array_equal(ea, eb) and np.array_equal(sa, sb) and np.array_equal(ua, ub)
    assert not np.array_equal(a[-1][2], a[0][2])


def test_batched_sweeps_replayed_as_graphs():
    """fans_solve_batch on a small grid: every sweep after the first is a graph replay over all lanes; bit-identical to plain launches,
    also when the cached graphs serve a second call and when a lane freezes between two replays"""
    ms = util.two_phase_ms(0, 7, (32, 32, 32))
    base = np.array(LOADS[0])
    macro = np.stack([base, 1e-3 * base, np.eye(6)[3] * 0.01])
    runs = {}
    for graph in (True, False):
        os.environ["FANS_GRAPH"] = "1" if graph else "0"
        try:
            ctx = simple.linear_elastic_context(ms, [1.0, 1.0, 1.0], BULK, SHEAR, "HEX8")
            first = ctx.solve_batch(macro, 200, 1e-9, "L2", "absolute")
            l0 = ctx.launch_count()
            second = ctx.solve_batch(macro[::-1].copy(), 200, 1e-9, "L2", "absolute")
            us = []
            for l in range(3):
                ctx.batch_displacement(l, "u_prev")
                us.append(ctx.download("u_prev"))
            runs[graph] = (first, second, us, ctx.launch_count() - l0)
            ctx.close()
        finally:
            os.environ.pop("FANS_GRAPH", None)
    for k in (0, 1):
        (ra, sa), (rb, sb) = runs[True][k], runs[False][k]
        assert [r["iters"] for r in ra] == [r["iters"] for r in rb]
        assert len(set(r["iters"] for r in ra)) > 1          # lanes froze at different sweeps
        assert all(np.array_equal(x["err_all"], y["err_all"]) for x, y in zip(ra, rb))
        assert np.array_equal(sa, sb)
    assert all(np.array_equal(a, b) for a, b in zip(runs[True][2], runs[False][2]))
    assert runs[True][3] == runs[False][3]
    # the second call solved the same lanes in reverse order
    assert np.array_equal(runs[True][0][1][::-1], runs[True][1][1])
