#!/usr/bin/env python
"""Nonlinear-path timing on one GPU: BASELINE config 3 at a size that fits one B200 (J2 plasticity with history on a synthetic fibre
image, CG with secant line search, incremental uniaxial strain): iterations, residual evaluations, ms per iteration and per kernel
class for every load step.  Usage: python tools/nlbench.py [--size 256] [--steps 4] [--de 1e-3]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fans_b200 import simple  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=256)
ap.add_argument("--steps", type=int, default=4)
ap.add_argument("--de", type=float, default=1e-3)
ap.add_argument("--fe", default="HEX8")
args = ap.parse_args()
n = args.size
ms = simple.fiber_microstructure(n)
ctx = simple.j2_fiber_context(ms, [1.0, 1.0, 1.0], args.fe, 0)
out = {"workload": "J2ViscoPlastic_NonLinearIsotropicHardening matrix + elastic fibres (vf %.3f), %d^3, %s, CG, eps11 in steps of %g"
                   % (float(ms.mean()), n, args.fe, args.de), "steps": []}
for t in range(args.steps):
    ctx.set_gradient([args.de * (t + 1), 0, 0, 0, 0, 0])
    ctx.set_profiling(True)
    r = ctx.solve("cg", 200, 1e-10, "Linfinity", "absolute")
    prof = ctx.profile()
    ctx.set_profiling(False)
    sig = ctx.homogenized_stress()
    out["steps"].append({"eps11": args.de * (t + 1), "iters": r["iters"], "residual_evals": r["n_residual_evals"], "loop_ms": r["loop_ms"],
                         "ms_per_iter": r["loop_ms"] / max(r["iters"], 1), "sigma11": float(sig[0]),
                         "kernel_ms": {k: round(v[0] / v[1], 4) for k, v in prof.items()},
                         "kernel_calls": {k: v[1] for k, v in prof.items()}})
    ctx.extrapolate_displacement()
print(json.dumps(out))
ctx.close()
