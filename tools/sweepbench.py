#!/usr/bin/env python
"""Device time of ONE element sweep (residual / homogenized stress) per material law — the nonlinear-path kernel k_sweep_sf (or
k_sweep with FANS_SWEEP_DENSE=1 for A/B runs).  Usage: python tools/sweepbench.py [--size 256] [--laws linear,j2,neohooke] [--reps 3]"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fans_b200 import simple, _lib as L  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=256)
ap.add_argument("--laws", default="linear,j2,neohooke")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--tag", default="")
args = ap.parse_args()
n = args.size
out = {"tag": args.tag, "size": n, "env": {k: v for k, v in os.environ.items() if k.startswith("FANS_")}, "laws": {}}
rng = np.random.default_rng(0)
for law in args.laws.split(","):
    if law == "linear":
        ms = simple.sphere_microstructure(n)
        ctx = simple.linear_elastic_context(ms, [1.0, 1.0, 1.0], [62.5, 222.222], [28.8462, 166.6667], "HEX8", 0)
        g0 = [0.001, -0.002, 0.003, 0.0015, -0.0025, 0.001]
    elif law == "j2":
        ms = simple.fiber_microstructure(n)
        ctx = simple.j2_fiber_context(ms, [1.0, 1.0, 1.0], "HEX8", 0)
        g0 = [0.004, 0, 0, 0, 0, 0]
    elif law == "neohooke":
        ms = simple.sphere_microstructure(n)
        ctx = simple.neohooke_context(ms, [1.0, 1.0, 1.0], [62.5, 222.222], [28.8462, 166.6667], "HEX8", 0)
        g0 = [1.0, 0, 0, 0, 1.0, 0, 0, 0, 1.1]
    else:
        raise SystemExit("unknown law " + law)
    ctx.set_gradient(g0)
    # a small smooth fluctuation so that the plastic / finite-strain branches see non-trivial input
    x = np.arange(n) / n
    u = np.zeros(ctx.field_shape)
    u[..., 0] = 1e-4 * np.sin(2 * np.pi * x)[:, None, None] * np.cos(2 * np.pi * x)[None, :, None]
    u[..., 1] = 1e-4 * np.sin(2 * np.pi * x)[None, None, :]
    ctx.upload("u", u)
    del u
    ctx.residual("r", "u")
    ctx.homogenized_stress()
    ctx.set_profiling(True)
    for _ in range(args.reps):
        ctx.residual("r", "u")
        sig = ctx.homogenized_stress()
    prof = ctx.profile()
    ctx.set_profiling(False)
    out["laws"][law] = {"kernel_ms": {k: round(v[0] / v[1], 4) for k, v in prof.items()}, "sigma": [float(s) for s in sig],
                        "r_linf": ctx.norm("r")}
    ctx.close()
print(json.dumps(out))
