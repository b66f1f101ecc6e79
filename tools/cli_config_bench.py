#!/usr/bin/env python
"""BASELINE.json configs 3 and 5 at a size that fits one B200, driven end to end through the C++ host front end (the drop-in path:
FANS JSON input -> FANS_gpu -> results directory), printing the solver's own 'Total Time per iteration' lines (include/solver.h:293-297).
Usage: python tools/cli_config_bench.py {j2|neohooke} [--size 256] [--steps 3]"""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cpp_host  # noqa: E402  (build helper of the C++ front end)
from fans_b200 import simple  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("config", choices=["j2", "neohooke"])
ap.add_argument("--size", type=int, default=256)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--method", default="cg")
args = ap.parse_args()
n = args.size
base = {"microstructure": {"filepath": "unused.h5", "datasetname": "/bench/ms", "L": [1.0, 1.0, 1.0]}, "FE_type": "HEX8", "method": args.method,
        "error_parameters": {"measure": "Linfinity", "type": "absolute", "tolerance": 1e-10}, "n_it": 500,
        "results": ["stress_average", "strain_average", "absolute_error"]}
if args.config == "j2":   # config 3: J2 matrix (test_J2Plasticity.json parameters) + elastic fibres along z, uniaxial strain ramp
    ms = simple.fiber_microstructure(n)
    cfg = dict(base, problem_type="mechanical", strain_type="small", materials=[
        {"phases": [0], "matmodel": "J2ViscoPlastic_NonLinearIsotropicHardening",
         "material_properties": {"bulk_modulus": [62.5], "shear_modulus": [28.8462], "yield_stress": [0.1], "isotropic_hardening_parameter": [0.0],
                                 "kinematic_hardening_parameter": [0.0], "viscosity": [1.0], "time_step": 0.01, "saturation_stress": [0.15], "saturation_exponent": [1000.0]}},
        {"phases": [1], "matmodel": "LinearElasticIsotropic", "material_properties": {"bulk_modulus": [222.222], "shear_modulus": [166.6667]}}],
        macroscale_loading=[[[1e-3 * (t + 1), 0, 0, 0, 0, 0] for t in range(args.steps)]])
else:                     # config 5: Neo-Hooke sphere, mixed BC of test_MixedBCs_LargeStrain.json load case 1 (F33 ramp, P11 = P22 = 0)
    ms = simple.sphere_microstructure(n)
    cfg = dict(base, problem_type="mechanical", strain_type="large", materials=[
        {"phases": [0, 1], "matmodel": "CompressibleNeoHookean",
         "material_properties": {"bulk_modulus": [62.5, 222.222], "shear_modulus": [28.8462, 166.6667]}}],
        macroscale_loading=[{"strain_indices": [1, 2, 3, 5, 6, 7, 8], "stress_indices": [0, 4],
                             "strain": [[0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0 + 0.1 * (t + 1)] for t in range(args.steps)],
                             "stress": [[0.0, 0.0] for _ in range(args.steps)]}])
exe = cpp_host.build()
with tempfile.TemporaryDirectory() as tmp:
    msf = os.path.join(tmp, "ms.u16")
    ms.tofile(msf)
    inp = os.path.join(tmp, "in.json")
    open(inp, "w").write(json.dumps(cfg))
    r = subprocess.run([exe, inp, os.path.join(tmp, "results"), msf, str(n), str(n), str(n)], capture_output=True, text=True)
    if r.returncode != 0:
        sys.exit(r.stdout[-2000:] + r.stderr[-2000:])
    its = re.findall(r"Total Time per iteration \.+ ([0-9.]+) sec", r.stdout)
    fft = re.findall(r"FFT Time per iteration \.+ ([0-9.]+) sec", r.stdout)
    tot = re.findall(r"Total Time \.+ ([0-9.]+) sec", r.stdout)
    eff = re.findall(r"# Effective Stress \.\. \(([^)]*)\)", r.stdout)
    print(json.dumps({"config": args.config, "size": n, "method": args.method, "load_steps": len(tot),
                      "total_time_per_iteration_s": [float(x) for x in its], "fft_time_per_iteration_s": [float(x) for x in fft],
                      "total_time_s": [float(x) for x in tot], "effective_stress_last": eff[-1].split() if eff else None}))
