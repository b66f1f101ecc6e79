"""CPU tests that PIN the oracle (oracle/fans_oracle.py) to numbers the reference itself produced.

The reference repository holds exactly two reference-computed known answers, both embedded in its test inputs
(SURVEY.md section 4 / 8c):
  KAT-1  test/input_files/test_MixedBCs.json:71-75        load case 2 prescribes sigma_bar = -0.05 / -0.1 (hydrostatic);
         load case 3 prescribes the strains the reference obtained for it.
  KAT-2  test/input_files/test_MixedBCs_LargeStrain.json:53-59   P_bar_33 = 74.757449712464 <-> F_bar = diag(0.8325.., 0.8325.., 2.0)
plus the property checks of test/pytest/*.py, restated here without MSUtils/HDF5.
"""
import numpy as np
import pytest

import fans_oracle as fo
import golden_util as gu


def test_sphere32_fixture_matches_formula():
    """sphere32.h5 decoded by tests/golden/make_golden.py == the analytic sphere used for the synthetic benches."""
    ms = gu.sphere32()
    assert ms.shape == (32, 32, 32) and int(ms.sum()) == 8744
    assert np.array_equal(ms, fo.sphere_microstructure(32))


def _run(name, lc, steps):
    cfg = gu.reference_input(name, [lc])
    res, sol = fo.run_load_cases(gu.sphere32(), cfg, max_steps=steps)
    return res[0], sol


def test_kat1_prescribed_strain_gives_reference_stress():
    """test_MixedBCs.json load case 3: the reference's strains -> its prescribed stresses -0.05, -0.1."""
    steps, _ = _run("MixedBCs", 2, 2)
    for st, target in zip(steps, (-0.05, -0.1)):
        s = st["stress_average"]
        assert np.allclose(s[:3], target, rtol=1e-9, atol=0.0), s
        assert np.abs(s[3:]).max() < 1e-12
    assert steps[0]["iters"] == 15  # 15 CG iterations


def test_kat1_prescribed_stress_gives_reference_strain():
    """test_MixedBCs.json load case 2 (pure stress control) -> the strains of load case 3 (mixed-BC update loop)."""
    steps, _ = _run("MixedBCs", 1, 1)
    g0 = steps[0]["g0"]
    assert np.allclose(g0[:3], -0.000201177817616389, rtol=1e-8, atol=0.0), g0
    assert np.abs(g0[3:]).max() < 1e-15
    assert np.allclose(steps[0]["stress_average"][:3], -0.05, rtol=1e-9)


def test_kat2_large_strain_reference_stress():
    """test_MixedBCs_LargeStrain.json load case 3: F_bar -> P_bar_33 = 74.757449712464 (compressible Neo-Hooke)."""
    steps, sol = _run("MixedBCs_LargeStrain", 2, 1)
    P = steps[0]["stress_average"]
    assert abs(P[8] / 74.757449712464 - 1.0) < 1e-10, P
    off = np.delete(P, 8)
    assert np.abs(off).max() < 1e-8, P  # reference: 3e-12 on the diagonal, 0 elsewhere
    assert steps[0]["iters"] == 57 and steps[0]["n_residual_evals"] == 161


@pytest.mark.parametrize("name", ["LinearThermal", "LinearElastic"])
def test_reference_pytest_invariants(name):
    """test/pytest/test_loading_to_strain_average.py:81, test_strain_stress_averaging.py:94-96,
    test_displacement_averaging.py:84-86, test_homogenization_consistency.py:94-96, test_homogenized_tangent_spd.py:79-95."""
    cfg = gu.reference_input(name)
    out = {}

    def on_step(sol, lc, t, res):
        out.update(sol.postprocess())
        out["tangent"] = sol.get_homogenized_tangent(1e-6)

    res, sol = fo.run_load_cases(gu.sphere32(), cfg, on_step=on_step)
    load = np.array(cfg["macroscale_loading"][0][0], dtype=float)
    assert np.allclose(out["strain_average"], load)
    n_str = sol.n_str
    assert np.allclose(out["strain"].reshape(-1, n_str).mean(0), out["strain_average"], rtol=1e-5, atol=1e-8)
    assert np.allclose(out["stress"].reshape(-1, n_str).mean(0), out["stress_average"], rtol=1e-5, atol=1e-8)
    assert np.allclose(out["displacement_fluctuation"].reshape(-1, sol.h).mean(0), 0.0, atol=1e-8)
    C = out["tangent"]
    assert np.allclose(C @ out["strain_average"], out["stress_average"], rtol=1e-4, atol=1e-1)
    assert np.allclose(C, C.T, rtol=1e-5, atol=1e-8)
    assert np.linalg.eigvalsh(C).min() > 0
    # the committed GPU-side fixture is what the oracle produces
    gold = gu.oracle_scenario(name)[0]
    assert res[0][0]["iters"] == gold["iters"]
    assert np.allclose(out["stress_average"], gold["stress_average"], rtol=1e-12, atol=1e-18)


def test_survey_anchor_values():
    """SURVEY.md section 6 sanity anchors of the restatement (thermal HEX8R: 16 its, elastic HEX8R: 18 its)."""
    g = gu.oracle_scenario("LinearThermal")[0]
    assert g["iters"] == 16 and abs(g["err_all"][0] - 1.7578125e-4) < 1e-15
    assert np.allclose(g["stress_average"], [0.017806793057, 0.035613586113, -0.017806793057], rtol=1e-9)
    g = gu.oracle_scenario("LinearElastic")[0]
    assert g["iters"] == 18 and abs(g["err_all"][0] - 9.400452473958e-4) < 1e-14
    assert np.allclose(g["stress_average"], [0.195889193692, -0.082381484564, 0.381402979196, 0.124021458809, -0.206702431348,
                                             0.082680972539], rtol=1e-9)


def test_golden_kats_recorded_in_fixture():
    """The fixture the GPU tests are compared with carries the two known answers."""
    m = {(r["load_case"], r["step"]): r for r in gu.oracle_scenario("MixedBCs")}
    assert np.allclose(m[(2, 0)]["stress_average"][:3], -0.05, rtol=1e-9)
    assert np.allclose(m[(2, 1)]["stress_average"][:3], -0.1, rtol=1e-9)
    assert np.allclose(m[(1, 0)]["g0"][:3], -0.000201177817616389, rtol=1e-8)
    m = {(r["load_case"], r["step"]): r for r in gu.oracle_scenario("MixedBCs_LargeStrain")}
    assert abs(m[(2, 0)]["stress_average"][8] / 74.757449712464 - 1) < 1e-10
    assert abs(m[(1, 0)]["g0"][0] / 0.832542244829 - 1) < 1e-9 and abs(m[(1, 0)]["g0"][8] / 2.000000000001 - 1) < 1e-9


def test_error_norm_is_max_over_slabs():
    """solver.h:430: Allreduce MAX for L1/L2 too => with P ranks the L2 'norm' is the max of per-slab norms."""
    ms = fo.sphere_microstructure(16)
    ep = {"measure": "L2", "type": "absolute", "tolerance": 1e-10}
    mats = [{"phases": [0, 1], "matmodel": "LinearThermalIsotropic", "material_properties": {"conductivity": [1.0, 10.0]}}]
    s1 = fo.OracleSolver(ms, [1, 1, 1], "thermal", mats, "HEX8", "cg", "small", ep, 3, n_ranks=1)
    s2 = fo.OracleSolver(ms, [1, 1, 1], "thermal", mats, "HEX8", "cg", "small", ep, 3, n_ranks=2)
    r = np.random.default_rng(0).standard_normal(s1.u.shape)
    e1, e2 = s1.compute_error(r), s2.compute_error(r)
    assert abs(e1 - np.sqrt((r * r).sum())) < 1e-12
    assert abs(e2 - max(np.sqrt((r[:8] ** 2).sum()), np.sqrt((r[8:] ** 2).sum()))) < 1e-12 and e2 < e1


def test_singular_green_blocks():
    """HEX8R has 80 singular Gamma blocks at 32^3 (hourglass modes + xi = 0), HEX8 only xi = 0."""
    mats = [{"phases": [0, 1], "matmodel": "LinearThermalIsotropic", "material_properties": {"conductivity": [1.0, 10.0]}}]
    ms = gu.sphere32()
    assert fo.OracleSolver(ms, [1, 1, 1], "thermal", mats, "HEX8R").n_singular == 80
    assert fo.OracleSolver(ms, [1, 1, 1], "thermal", mats, "HEX8").n_singular == 1
