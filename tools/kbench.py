#!/usr/bin/env python
"""Per-kernel device times of the linear CG iteration (bench.py's workload) without the e2e / CPU legs — for quick A/B runs of
kernel variants selected through environment variables.  Usage: python tools/kbench.py [--size 512] [--steps 5] [--tag name]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fans_b200 import simple  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=512)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--tag", default="")
ap.add_argument("--fe", default="HEX8")
ap.add_argument("--ms", default="ellipsoid", help="ellipsoid | homogeneous | layers | voronoi | poly16 (16 rotated cubic grains, one tensor each)")
ap.add_argument("--no-profile", action="store_true", help="only the timed solve (used under ncu by bench.py)")
args = ap.parse_args()
n = args.size
dims = [n, n, n]
ms = simple.ellipsoid_microstructure(dims)
if args.ms == "homogeneous":
    ms[...] = 0
    ms[0, 0, 0] = 1
elif args.ms == "voronoi":
    ms = simple.voronoi_microstructure(dims)
elif args.ms == "layers":
    ms[...] = 0
    ms[n // 4: 3 * n // 4] = 1
if args.ms == "poly16":   # more phases than the constant-bank stencil holds: the coefficient-table instantiation (stencil.cu, NQ = 0)
    ms = simple.voronoi_labels(dims, 16)
    ctx = simple.linear_elastic_tensor_context(ms, [1.0, 1.0, 1.0], simple.rotated_cubic_tangents(16), args.fe, 0)
else:
    ctx = simple.linear_elastic_context(ms, [1.0, 1.0, 1.0], [62.5, 222.222], [28.8462, 166.6667], args.fe, 0)
ctx.set_gradient([0.001, -0.002, 0.003, 0.0015, -0.0025, 0.001])
ctx.solve("cg", 3, 0.0, "Linfinity", "absolute")
ctx.zero("u")
res = ctx.solve("cg", args.steps, 0.0, "Linfinity", "absolute")
loop = res["loop_ms"] / args.steps
if args.no_profile:
    print(json.dumps({"tag": args.tag, "ms_per_iter": round(loop, 4)}))
    ctx.close()
    sys.exit(0)
ctx.zero("u")
ctx.set_profiling(True)
ctx.solve("cg", args.steps, 0.0, "Linfinity", "absolute")
prof = ctx.profile()
sig = ctx.homogenized_stress()
out = {"tag": args.tag, "ms_per_iter": round(loop, 4), "kernel_ms": {k: round(v[0] / v[1], 4) for k, v in prof.items()},
       "sigma": [float(x) for x in sig], "env": {k: v for k, v in os.environ.items() if k.startswith("FANS_")}}
print(json.dumps(out))
ctx.close()
