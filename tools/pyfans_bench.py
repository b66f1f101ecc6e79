#!/usr/bin/env python
"""One micro simulation of the preCICE Micro Manager use case (pyfans/micro.cpp:40-85): PyFANS.MicroSimulation.solve = solve +
homogenized stress + homogenized tangent, repeated on the same microstructure with changing macroscopic strains.  Wall time per call.
  python tools/pyfans_bench.py [n=32] [calls=20]      (FANS_TANGENT_BATCH=0 / FANS_GRAPH=0 in the environment for the A/B numbers)"""
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402
from fans_b200 import simple  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 20
g.build_pyfans()
sys.path.insert(0, os.path.join(ROOT, "fans_b200", "lib"))
import PyFANS  # noqa: E402

tmp = tempfile.mkdtemp()
ms = simple.sphere_microstructure(n)
np.save(os.path.join(tmp, "ms.npy"), np.ascontiguousarray(ms.transpose(2, 1, 0)).astype(np.uint8))
cfg = {"microstructure": {"filepath": "ms.npy", "datasetname": "/ms", "L": [1.0, 1.0, 1.0]}, "problem_type": "mechanical", "strain_type": "small",
       "materials": [{"phases": [0, 1], "matmodel": "LinearElasticIsotropic",
                      "material_properties": {"bulk_modulus": [62.5, 222.222], "shear_modulus": [28.8462, 166.6667]}}],
       "FE_type": "HEX8", "method": "cg", "error_parameters": {"measure": "Linfinity", "type": "absolute", "tolerance": 1e-10}, "n_it": 100,
       "macroscale_loading": [[[0, 0, 0, 0, 0, 0]]], "results": []}
open(os.path.join(tmp, "input.json"), "w").write(json.dumps(cfg))
os.chdir(tmp)
sim = PyFANS.MicroSimulation(0)
rng = np.random.default_rng(0)
strain = np.array([0.001, -0.002, 0.003, 0.0015, -0.0025, 0.001])
out = sim.solve({"strains1to3": strain[:3], "strains4to6": strain[3:]}, 0.1)   # warm-up: allocations, graphs
t0 = time.perf_counter()
for i in range(calls):
    e = strain * (1.0 + 0.1 * rng.standard_normal(6))
    out = sim.solve({"strains1to3": e[:3], "strains4to6": e[3:]}, 0.1)
dt = (time.perf_counter() - t0) / calls
print(json.dumps({"grid": n, "calls": calls, "ms_per_micro_solve": round(1e3 * dt, 3), "cmat1": [float(x) for x in out["cmat1"]],
                  "env": {k: v for k, v in os.environ.items() if k.startswith("FANS_")}}))
