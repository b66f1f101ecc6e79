"""The C++ host front end end to end on the GPU: FANS_gpu <input.json> <results_dir> on reference scenarios, results read back
from the sink and compared with the committed oracle fixture (the same numbers the reference's pytest suite looks at)."""
import json
import os
import subprocess

import numpy as np
import pytest

import golden_util as gu
import cpp_host
from util import rel_err

pytestmark = pytest.mark.gpu


def run_cli(tmp_path, cfg, env=None):
    exe = cpp_host.build()
    ms = tmp_path / "ms.u16"
    gu.sphere32().tofile(ms)
    inp = tmp_path / "in.json"
    inp.write_text(json.dumps(cfg))
    out = tmp_path / "results"
    r = subprocess.run([exe, str(inp), str(out), str(ms), "32", "32", "32"], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, **(env or {})))
    assert r.returncode == 0, r.stderr[-2000:]
    index = [json.loads(l) for l in open(out / "index.jsonl")]
    def load(name, lc, t):
        e = [i for i in index if i["name"] == name and i["load"] == lc and i["time_step"] == t][0]
        dt = {"f64": np.float64, "f32": np.float32, "u16": np.uint16}[e["dtype"]]
        return np.fromfile(str(out) + e["path"], dtype=dt).reshape(e["dims"])
    return load, r.stdout


@pytest.mark.parametrize("name,steps", [("LinearElastic", 1), ("LinearThermal", 1), ("MixedBCs", 2)])
def test_cli_scenario(tmp_path, name, steps):
    cfg = gu.reference_input(name)
    for i, lc in enumerate(cfg["macroscale_loading"]):   # bound the run like the fixture
        if isinstance(lc, dict):
            for k in ("strain", "stress"):
                lc[k] = lc[k][:steps]
        else:
            cfg["macroscale_loading"][i] = lc[:steps]
    load, stdout = run_cli(tmp_path, cfg)
    for g in gu.oracle_scenario(name):
        if g["step"] >= steps:
            continue
        sa = load("stress_average", g["load_case"], g["step"])
        assert rel_err(sa, g["stress_average"]) < 1e-9 or np.abs(np.array(g["stress_average"])).max() < 1e-12
        ea = load("strain_average", g["load_case"], g["step"])
        assert rel_err(ea, g["g0"]) < 1e-9
        err = load("absolute_error", g["load_case"], g["step"])
        assert abs(len(err) - 1 - g["iters"]) <= 1
        # reference pytest invariants on the written fields (test_strain_stress_averaging.py, test_displacement_averaging.py)
        stress = load("stress", g["load_case"], g["step"])
        assert np.allclose(stress.reshape(32 ** 3, -1).mean(0), sa, rtol=1e-5, atol=1e-8)
        uf = load("displacement_fluctuation", g["load_case"], g["step"])
        assert np.allclose(uf.reshape(32 ** 3, -1).mean(0), 0.0, atol=1e-8)
    if name in ("LinearElastic", "LinearThermal"):   # results: homogenized_tangent requested (test_homogenization_consistency.py)
        C = load("homogenized_tangent", 0, 0)
        sa, ea = load("stress_average", 0, 0), load("strain_average", 0, 0)
        assert np.allclose(C @ ea, sa, rtol=1e-4, atol=1e-1) and np.allclose(C, C.T, rtol=1e-5, atol=1e-8)
        assert np.linalg.eigvalsh(C).min() > 0


@pytest.mark.parametrize("name", ["LinearElastic", "LinearThermal"])
def test_cli_tangent_batched_equals_the_reference_loop(tmp_path, name):
    """get_homogenized_tangent (solver.h:739-778): the batched form (one CG loop over the n_str unit load cases) against the
    one-solve-per-column loop of the reference, through the CLI; two time steps, so the solver state the tangent leaves behind
    (gradient, displacement, error type — quirks the reference has) feeds the second step identically in both forms"""
    cfg = gu.reference_input(name)
    first = cfg["macroscale_loading"][0][0]
    cfg["macroscale_loading"] = [[first, [2.0 * v for v in first]]]
    (tmp_path / "a").mkdir()
    (tmp_path / "b").mkdir()
    load_b, out_b = run_cli(tmp_path / "a", cfg)
    load_s, out_s = run_cli(tmp_path / "b", cfg, {"FANS_TANGENT_BATCH": "0"})
    assert "one batched CG loop" in out_b and "one batched CG loop" not in out_s
    for t in range(2):
        Cb, Cs = load_b("homogenized_tangent", 0, t), load_s("homogenized_tangent", 0, t)
        assert rel_err(Cb, Cs) < 1e-6      # both stop at the relative 1e-6 tolerance the tangent solves use
        assert rel_err(load_b("stress_average", 0, t), load_s("stress_average", 0, t)) < 1e-6
        # iteration counts of the step that FOLLOWS a tangent computation: same start state in both forms
        assert abs(len(load_b("absolute_error", 0, t)) - len(load_s("absolute_error", 0, t))) <= 1


def test_cli_hdf5_results(tmp_path):
    """FANS_gpu <input.json> results.h5: the reference's results file (include/reader.h:173-351) — same numbers as the directory sink,
    fields stored [Z][Y][X][extra] with permute_order = "zyx"."""
    import h5_minireader as h5
    cfg = gu.reference_input("LinearElastic")
    cfg["macroscale_loading"] = [lc[:1] for lc in cfg["macroscale_loading"]]
    exe = cpp_host.build()
    ms = tmp_path / "ms.u16"
    gu.sphere32().tofile(ms)
    inp = tmp_path / "in.json"
    inp.write_text(json.dumps(cfg))
    out = tmp_path / "results.h5"
    r = subprocess.run([exe, str(inp), str(out), str(ms), "32", "32", "32"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    w = h5.H5File(str(out)).walk()
    g = [x for x in gu.oracle_scenario("LinearElastic") if x["step"] == 0][0]
    pre = [k for k in w if k.endswith("/load%d/time_step0/stress_average" % g["load_case"])][0].rsplit("/", 1)[0]
    assert "_results/" in pre
    assert rel_err(w[pre + "/stress_average"][0], g["stress_average"]) < 1e-9
    stress, attrs = w[pre + "/stress"]
    assert attrs == {"permute_order": "zyx"} and stress.shape == (32, 32, 32, 6)
    assert np.allclose(stress.reshape(-1, 6).mean(0), w[pre + "/stress_average"][0], rtol=1e-5, atol=1e-8)
    msr, _ = w[pre + "/microstructure"]
    assert np.array_equal(msr[..., 0], np.transpose(gu.sphere32(), (2, 1, 0)))


def j2_cfg(steps):
    cfg = gu.reference_input("J2Plasticity")
    cfg["macroscale_loading"] = [[[0.002 * (t + 1), -0.001 * (t + 1), -0.001 * (t + 1), 0.0005 * (t + 1), 0, 0] for t in range(steps)]]
    cfg["results"] = ["stress_average", "strain_average", "absolute_error", "stress", "strain", "plastic_strain", "isotropic_hardening_variable"]
    return cfg


def test_cli_j2_multistep_matches_oracle(tmp_path):
    """J2 plasticity over several time steps through FANS_gpu, results requested like a user would (stress AND strain fields): the
    postprocess sweep must run exactly once per time step — J2Plasticity accumulates psi / psi_bar on every get_sigma call
    (J2Plasticity.h:103-104), so a second sweep would change every later step."""
    import fans_oracle as fo
    steps = 3
    cfg = j2_cfg(steps)
    load, _ = run_cli(tmp_path, cfg)
    out = []

    def o_step(sol, lc, t, res):
        res["post"] = sol.postprocess()   # one getStrainStress sweep per step, like Solver::postprocess

    ro, _ = fo.run_load_cases(gu.sphere32(), cfg, on_step=o_step)
    for t in range(steps):
        sa = load("stress_average", 0, t)
        assert rel_err(sa, ro[0][t]["post"]["stress_average"]) < 1e-9, t
        assert abs(len(load("absolute_error", 0, t)) - 1 - ro[0][t]["iters"]) <= 1
        assert rel_err(load("stress", 0, t), np.asarray(ro[0][t]["post"]["stress"]).reshape(32, 32, 32, 6)) < 1e-8, t
        assert rel_err(load("isotropic_hardening_variable", 0, t),
                       np.asarray(ro[0][t]["post"]["isotropic_hardening_variable"]).reshape(32, 32, 32)) < 1e-8, t


def test_cli_two_ranks(tmp_path):
    """FANS_gpu on 2 ranks (one per GPU, NCCL id through FANS_COMM_FILE, no MPI): same averages and iteration counts as one rank,
    the field slabs of the two ranks concatenate to the single-rank field.  Reference: mpiexec -n 2 FANS (src/main.cpp:60-61)."""
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cfg = j2_cfg(2)
    exe = cpp_host.build()
    ms = tmp_path / "ms.u16"
    gu.sphere32().tofile(ms)
    inp = tmp_path / "in.json"
    inp.write_text(json.dumps(cfg))
    one, two = tmp_path / "one", tmp_path / "two"
    r = subprocess.run([exe, str(inp), str(one), str(ms), "32", "32", "32"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "fans_mprun.py"), "-n", "2", "--", exe, str(inp), str(two), str(ms), "32", "32", "32"],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]

    def loader(d):
        index = [json.loads(l) for l in open(os.path.join(d, "index.jsonl"))]

        def load(name, t):
            e = [i for i in index if i["name"] == name and i["time_step"] == t][0]
            return np.fromfile(str(d) + e["path"], dtype=np.float64).reshape(e["dims"])
        return load
    l1, l2a, l2b = loader(one), loader(two), loader(two / "rank1")
    for t in range(2):
        assert rel_err(l2a("stress_average", t), l1("stress_average", t)) < 1e-9
        assert abs(len(l2a("absolute_error", t)) - len(l1("absolute_error", t))) <= 1
        both = np.concatenate([l2a("stress", t), l2b("stress", t)], axis=0)
        assert both.shape == (32, 32, 32, 6) and rel_err(both, l1("stress", t)) < 1e-8


def test_cli_gbdiffusion(tmp_path):
    """GBDiffusion end to end: image + attributes from an HDF5 file, 4 tags (2 crystals, 2 boundary phases with oblique normals), solved
    on the GPU through FANS_gpu; the oracle solves the same problem as LinearThermalTriclinic with the equivalent tensors."""
    import fans_oracle as fo
    import test_host_cpp as th
    exe = cpp_host.build()
    f = tmp_path / "gb.h5"
    assert subprocess.run([exe, "--gbselftest", str(f)]).returncode == 0
    cfg = th.gb_cfg(str(f))
    inp = tmp_path / "in.json"
    inp.write_text(json.dumps(cfg))
    out = tmp_path / "res"
    r = subprocess.run([exe, str(inp), str(out)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    index = [json.loads(l) for l in open(out / "index.jsonl")]
    e = [i for i in index if i["name"] == "stress_average"][0]
    q = np.fromfile(str(out) + e["path"], dtype=np.float64)
    kap = th.gb_expected()
    keys = {"K_11": (0, 0), "K_12": (0, 1), "K_13": (0, 2), "K_22": (1, 1), "K_23": (1, 2), "K_33": (2, 2)}
    mats = [{"phases": [0, 1, 2, 3], "matmodel": "LinearThermalTriclinic",
             "material_properties": {k: [float(t[ij]) for t in kap] for k, ij in keys.items()}}]
    ms = np.zeros((8, 4, 4), dtype=np.uint16)
    ms[3], ms[4:7], ms[7] = 2, 1, 3
    sol = fo.OracleSolver(ms, [2.0, 1.0, 1.0], "thermal", mats, "HEX8", "cg", "small", cfg["error_parameters"], 100)
    sol.set_gradient([0.01, 0.02, -0.01])
    sol.solve()
    assert rel_err(q, sol.get_homogenized_stress()) < 1e-9
    g = [i for i in index if i["name"] == "GBnormals"][0]
    nf = np.fromfile(str(out) + g["path"], dtype=np.float64).reshape(8, 4, 4, 3)
    assert np.allclose(nf[7, 0, 0], [0.6, 0.8, 0.0]) and np.allclose(nf[3, 1, 2], [1.0, 0.0, 0.0]) and not nf[0].any()
