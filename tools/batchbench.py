#!/usr/bin/env python
"""Six unit load cases of a two-phase linear-elastic image (the homogenized-tangent loop, solver.h:739-778): fans_solve_batch against
six fans_solve calls, wall clock and device time, for a range of grid sizes.  python tools/batchbench.py [sizes...]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fans_b200 import simple  # noqa: E402

K_BULK, G_SHEAR = [62.5, 222.222], [28.8462, 166.6667]


def run(n, reps):
    ms = simple.ellipsoid_microstructure((n, n, n))
    ctx = simple.linear_elastic_context(ms, [1.0, 1.0, 1.0], K_BULK, G_SHEAR, "HEX8")
    macro = 1e-3 * np.eye(6)

    def seq():
        its, sig, dev = 0, np.zeros((6, 6)), 0.0
        for i in range(6):
            ctx.zero("u")
            ctx.set_gradient(macro[i])
            r = ctx.solve("cg", 500, 1e-10, "Linfinity", "absolute")
            sig[i] = ctx.homogenized_stress()
            its += r["iters"]
            dev += r["loop_ms"]
        return its, sig, dev

    def bat():
        res, sig = ctx.solve_batch(macro, 500, 1e-10, "Linfinity", "absolute")
        return sum(r["iters"] for r in res), sig, res[0]["loop_ms"]

    out = {}
    for name, fn in (("sequential", seq), ("batched", bat)):
        fn()
        t0 = time.perf_counter()
        for _ in range(reps):
            its, sig, dev = fn()
        out[name] = ((time.perf_counter() - t0) / reps, its, sig, dev)
    ctx.close()
    ws, wb = out["sequential"][0], out["batched"][0]
    diff = np.abs(out["sequential"][2] - out["batched"][2]).max() / np.abs(out["sequential"][2]).max()
    print("%4d^3  sequential %9.3f ms wall (%3d its, loop %8.3f ms, %.4f ms/it)   batched %9.3f ms wall (%3d its, loop %8.3f ms, %.4f ms/it)   "
          "speed-up wall %.2fx loop %.2fx   sigma rel diff %.1e"
          % (n, 1e3 * ws, out["sequential"][1], out["sequential"][3], out["sequential"][3] / out["sequential"][1], 1e3 * wb, out["batched"][1],
             out["batched"][3], out["batched"][3] / out["batched"][1], ws / wb, out["sequential"][3] / out["batched"][3], diff), flush=True)


if __name__ == "__main__":
    sizes = [int(a) for a in sys.argv[1:]] or [16, 32, 64, 128, 256]
    for n in sizes:
        run(n, 20 if n <= 64 else (5 if n <= 128 else 2))
