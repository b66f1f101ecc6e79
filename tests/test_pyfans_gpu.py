"""PyFANS.MicroSimulation (pyfans/micro.cpp:27-95) on the GPU: one micro solve for the preCICE Micro Manager = solve +
homogenized stress + homogenized tangent (6 more solves), compared with the oracle."""
import json
import os
import sys

import numpy as np
import pytest

import fans_oracle as fo
import util
from util import rel_err

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_micro_simulation(tmp_path, monkeypatch):
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    g.build_pyfans()
    sys.path.insert(0, os.path.join(ROOT, "fans_b200", "lib"))
    import PyFANS
    ms = util.two_phase_ms(0, 21, (16, 16, 16))
    np.save(tmp_path / "ms.npy", np.ascontiguousarray(ms.transpose(2, 1, 0)).astype(np.uint8))   # on-disk order z-y-x like the HDF5 files
    cfg = {"microstructure": {"filepath": "ms.npy", "datasetname": "/ms", "L": [1.0, 1.0, 1.0]}, "problem_type": "mechanical",
           "strain_type": "small", "materials": util.ELASTIC, "FE_type": "HEX8", "method": "cg",
           "error_parameters": {"measure": "Linfinity", "type": "absolute", "tolerance": 1e-10}, "n_it": 100,
           "macroscale_loading": [[[0, 0, 0, 0, 0, 0]]], "results": []}
    (tmp_path / "input.json").write_text(json.dumps(cfg))
    monkeypatch.chdir(tmp_path)   # the reference hard-wires "input.json" in the working directory (micro.hpp:22)
    sim = PyFANS.MicroSimulation(0)
    strain = np.array([0.001, -0.002, 0.003, 0.0015, -0.0025, 0.001])
    out = sim.solve({"strains1to3": strain[:3], "strains4to6": strain[3:]}, 0.1)
    sol = fo.OracleSolver(ms, [1.0, 1.0, 1.0], "mechanical", util.ELASTIC, "HEX8", "cg", "small", cfg["error_parameters"], 100)
    sol.set_gradient(strain)
    sol.solve()
    sig = sol.get_homogenized_stress()
    C = sol.get_homogenized_tangent(1e-6)
    assert rel_err(np.concatenate([out["stresses1to3"], out["stresses4to6"]]), sig) < 1e-9
    tri = np.concatenate([out["cmat%d" % k] for k in range(1, 8)])
    assert rel_err(tri, C[np.triu_indices(6)]) < 1e-6   # the tangent solves stop at the RELATIVE 1e-6 tolerance (solver.h:749-750)
    # a second call keeps working on the same microstructure (Micro Manager calls solve() every coupling iteration)
    out2 = sim.solve({"strains1to3": 2 * strain[:3], "strains4to6": 2 * strain[3:]}, 0.1)
    assert np.allclose(out2["cmat1"], out["cmat1"], rtol=1e-5)
