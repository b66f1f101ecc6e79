#!/bin/bash
# compute-sanitizer over the code paths added late in round 2 (small grids, bounded): lanes of a batched solve, graph replay, the opt-in
# TMA z pass; memcheck everywhere, racecheck on one batched solve and the TMA pass
mkdir -p gpurun_out
cat > /tmp/san_case.py <<'PY'
import os, sys
import numpy as np
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests")); sys.path.insert(0, os.path.join(os.getcwd(), "oracle"))
import util
from fans_b200 import simple
shape = tuple(int(x) for x in sys.argv[1].split(","))
ms = util.two_phase_ms(0, 3, shape)
ctx = simple.linear_elastic_context(ms, [1.0, 1.0, 1.0], [62.5, 222.222], [28.8462, 166.6667], "HEX8")
res, sig = ctx.solve_batch(np.eye(6)[:3] * 0.01, 30, 1e-9, "Linfinity", "absolute")
ctx.set_gradient([0.001, -0.002, 0.003, 0.0015, -0.0025, 0.001])
r = ctx.solve("cg", 30, 1e-9, "Linfinity", "absolute")
r2 = ctx.solve("cg", 30, 1e-12, "Linfinity", "absolute")
print("ok", [x["iters"] for x in res], r["iters"], r2["iters"], float(np.abs(sig).max()), ctx.homogenized_stress()[:2])
ctx.close()
PY
{
echo "== memcheck 16x16x32 (graphs on)"; FANS_GRAPH=1 timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python /tmp/san_case.py 16,16,32 2>&1 | tail -6
echo "== memcheck 16x16x32 TMA z pass"; FANS_Z_TMA=1 FANS_GRAPH=0 timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python /tmp/san_case.py 16,16,32 2>&1 | tail -6
echo "== racecheck 8x16x16"; FANS_GRAPH=0 timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python /tmp/san_case.py 8,16,16 2>&1 | tail -6
echo "== racecheck 8x16x16 TMA z pass"; FANS_Z_TMA=1 FANS_GRAPH=0 timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python /tmp/san_case.py 8,16,16 2>&1 | tail -6
} > gpurun_out/r2sn_sanitizer.txt 2>&1
cat gpurun_out/r2sn_sanitizer.txt
