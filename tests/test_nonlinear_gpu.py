"""GPU parity of the nonlinear paths against the live oracle on small grids: every material law with its history
variables in the PLASTIC / finite-strain regime, FE types, the fixed-point solver, secant line search."""
import numpy as np
import pytest

import fans_oracle as fo
import util
from util import ELASTIC, rel_err

pytestmark = pytest.mark.gpu

EL1 = {"phases": [1], "matmodel": "LinearElasticIsotropic", "material_properties": {"bulk_modulus": [222.222], "shear_modulus": [166.6667]}}
BASE = {"bulk_modulus": [62.5], "shear_modulus": [28.8462], "yield_stress": [0.1]}
MODELS = {
    "pp_lin": {"phases": [0], "matmodel": "PseudoPlasticLinearHardening", "material_properties": dict(BASE, hardening_parameter=[5.0])},
    "pp_nonlin": {"phases": [0], "matmodel": "PseudoPlasticNonLinearHardening",
                  "material_properties": dict(BASE, hardening_exponent=[0.2], eps_0=[0.01])},
    "j2_lin": {"phases": [0], "matmodel": "J2ViscoPlastic_LinearIsotropicHardening",
               "material_properties": dict(BASE, isotropic_hardening_parameter=[3.0], kinematic_hardening_parameter=[2.0], viscosity=[1.0],
                                           time_step=0.01)},
    "j2_nonlin": {"phases": [0], "matmodel": "J2ViscoPlastic_NonLinearIsotropicHardening",
                  "material_properties": dict(BASE, isotropic_hardening_parameter=[0.0], kinematic_hardening_parameter=[0.0], viscosity=[1.0],
                                              time_step=0.01, saturation_stress=[0.15], saturation_exponent=[1000.0])},
    "j2new": {"phases": [0], "matmodel": "J2PlasticityNew_LinearIsotropicHardening",
              "material_properties": dict(BASE, isotropic_hardening_parameter=[4.0])},
}
LOAD = [[0.002, -0.001, -0.001, 0.0005, 0, 0], [0.004, -0.002, -0.002, 0.001, 0, 0], [0.003, -0.0015, -0.0015, 0.0005, 0, 0]]


def cfg_for(mats, fe, method, loading, problem="mechanical", strain_type="small", n_it=200, tol=1e-10):
    return {"microstructure": {"L": [1.0, 1.5, 2.0]}, "problem_type": problem, "strain_type": strain_type, "materials": mats, "FE_type": fe,
            "method": method, "error_parameters": {"measure": "Linfinity", "type": "absolute", "tolerance": tol}, "n_it": n_it,
            "macroscale_loading": [loading]}


def compare(cfg, ms, check_hist=()):
    o_out, g_out = [], []

    def o_step(sol, lc, t, res):
        res["stress_average"] = sol.get_homogenized_stress()
        res["u"] = sol.u.copy()
        res["post"] = sol.postprocess()

    def g_step(ctx, lc, t, res):
        res["stress_average"] = ctx.homogenized_stress()
        res["u"] = ctx.download("u")
        res["post"] = {}
        # ONE sweep like Solver::postprocess (solver.h:497-545): element averages and every Gauss point (strain_gp / stress_gp)
        res["post"]["strain"], res["post"]["stress"], res["post"]["strain_gp"], res["post"]["stress_gp"] = ctx.strain_stress_gp()
        res["post"].update({k: ctx.get_field(k) for k in check_hist})

    ro, sol = fo.run_load_cases(ms, cfg, on_step=o_step)
    rg, ctx = util.run_gpu_load_cases(ms, cfg, on_step=g_step)
    for t, (a, b) in enumerate(zip(rg[0], ro[0])):
        assert abs(a["iters"] - b["iters"]) <= 1, (t, a["iters"], b["iters"])
        assert rel_err(a["stress_average"], b["stress_average"]) < 1e-9, (t, a["stress_average"], b["stress_average"])
        assert rel_err(a["g0"], b["g0"]) < 1e-9
        if cfg["FE_type"] != "HEX8R":
            # HEX8R carries zero-energy hourglass modes: Gamma blocks whose noise-level singular values straddle the reference's
            # ABSOLUTE 1e-14 cut (solver.h:189-191) are decided by rounding, so u is only defined up to those modes there;
            # strain and stress (compared below) do not see them.
            assert rel_err(a["u"], b["u"]) < 1e-8, t
        for k in a["post"]:
            assert rel_err(a["post"][k], np.asarray(b["post"][k]).reshape(a["post"][k].shape)) < 1e-8, (t, k)
    ctx.close()
    return rg, ro


@pytest.mark.parametrize("model,hist", [("pp_lin", ("plastic_flag",)), ("pp_nonlin", ("plastic_flag",)),
                                        ("j2_lin", ("plastic_strain", "isotropic_hardening_variable", "kinematic_hardening_variable")),
                                        ("j2_nonlin", ("plastic_strain", "isotropic_hardening_variable", "kinematic_hardening_variable")),
                                        ("j2new", ("plastic_strain", "isotropic_hardening_variable"))])
@pytest.mark.parametrize("fe", ["HEX8", "BBAR", "HEX8R"])
def test_small_strain_plasticity(model, hist, fe):
    if fe != "HEX8" and model in ("pp_lin", "j2_lin"):
        pytest.skip("FE variants are exercised with one model per family")
    ms = util.two_phase_ms(0, 11, (16, 8, 32))
    rg, ro = compare(cfg_for([MODELS[model], EL1], fe, "cg", LOAD), ms, hist)
    assert ro[0][0]["n_residual_evals"] > 5  # the nonlinear branch really ran


@pytest.mark.parametrize("model", ["CompressibleNeoHookean", "SaintVenantKirchhoff"])
@pytest.mark.parametrize("method", ["cg", "fp"])
def test_large_strain(model, method):
    ms = util.two_phase_ms(0, 12, (8, 16, 16))
    mats = [{"phases": [0, 1], "matmodel": model, "material_properties": {"bulk_modulus": [62.5, 222.222], "shear_modulus": [28.8462, 166.6667]}}]
    load = [[1.0, 0.05, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.02], [1.0, 0.1, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.04]]
    cfg = cfg_for(mats, "HEX8", method, load, strain_type="large", n_it=400, tol=1e-9)
    if method == "fp":
        # With the reference's own large-strain reference medium (P<=Q sum, shear entries mu/2: LargeStrainMechModel.h:142-171) the
        # basic scheme diverges in the reference as well; give it a stiff user "reference_material" (MaterialManager.h:179-196).
        lam, mu = 222.222 - 2.0 / 3.0 * 166.6667, 166.6667
        C = np.zeros((9, 9))
        for i in range(3):
            for j in range(3):
                C[3 * i + i, 3 * j + j] += lam
                C[3 * i + j, 3 * i + j] += mu
        cfg["reference_material"] = C.tolist()
        cfg["macroscale_loading"] = [[[1.0, 0.02, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.01], [1.0, 0.04, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.02]]]
    compare(cfg, ms)


def test_fixed_point_linear_elastic():
    ms = util.two_phase_ms(0, 13, (16, 16, 16))
    compare(cfg_for(ELASTIC, "HEX8", "fp", [[0.001, -0.002, 0.003, 0.0015, -0.0025, 0.001]], n_it=300, tol=1e-9), ms)


def test_mixed_bc_small_strain_plastic():
    ms = util.two_phase_ms(0, 14, (16, 16, 16))
    load = {"strain_indices": [2, 3, 4, 5], "stress_indices": [0, 1], "strain": [[0.002, 0, 0, 0], [0.004, 0, 0, 0]], "stress": [[0, 0], [0.01, 0]]}
    compare(cfg_for([MODELS["pp_nonlin"], EL1], "HEX8", "cg", load, n_it=300), ms)


def test_neg_jacobian_is_reported():
    """CompressibleNeoHookean.h:40-42 throws on J <= 0; the device raises a sticky fault surfaced as FANS_ERR_NEG_JACOBIAN."""
    from fans_b200 import _lib as L
    ms = util.two_phase_ms(0, 15, (8, 8, 8))
    mats = [{"phases": [0, 1], "matmodel": "CompressibleNeoHookean", "material_properties": {"bulk_modulus": [62.5, 222.222], "shear_modulus": [28.8462, 166.6667]}}]
    par = fo.OracleSolver(ms, [1, 1, 1], "mechanical", mats, "HEX8", "cg", "large", None, 0)
    ctx = util.ctx_from_oracle(par)
    ctx.set_gradient([-1.0, 0, 0, 0, 1.0, 0, 0, 0, 1.0])
    with pytest.raises(L.FansError, match="Negative Jacobian"):
        ctx.residual("r", "u")
    ctx.close()


def test_history_gauss_point_outputs():
    """plastic_strain_gp / isotropic_hardening_variable_gp / kinematic_hardening_variable_gp (J2Plasticity.h:279-307): every Gauss point of
    the committed state, [element][gp][component]; their Gauss-point means are the element outputs, and they match the oracle's state."""
    ms = util.two_phase_ms(0, 11, (16, 8, 32))
    cfg = cfg_for([MODELS["j2_lin"], EL1], "HEX8", "cg", LOAD)
    sol = fo.OracleSolver(ms, cfg["microstructure"]["L"], cfg["problem_type"], cfg["materials"], cfg["FE_type"], cfg["method"], "small",
                          cfg["error_parameters"], cfg["n_it"])
    ctx = util.ctx_from_oracle(sol)
    ep = cfg["error_parameters"]
    for g in cfg["macroscale_loading"][0]:
        sol.set_gradient(g)
        ctx.set_gradient(g)
        sol.solve()
        ctx.solve("cg", cfg["n_it"], ep["tolerance"], ep["measure"], ep["type"])
    eg = ctx.get_field("plastic_strain_gp")
    assert eg.shape == ms.shape + (8, 6) and np.abs(eg).max() > 1e-6          # the plastic branch really ran
    assert rel_err(eg.mean(-2), ctx.get_field("plastic_strain")) < 1e-12
    assert rel_err(ctx.get_field("isotropic_hardening_variable_gp").mean(-1), ctx.get_field("isotropic_hardening_variable")) < 1e-12
    assert rel_err(ctx.get_field("kinematic_hardening_variable_gp").mean(-2), ctx.get_field("kinematic_hardening_variable")) < 1e-12
    m = sol.models[0]
    assert rel_err(eg, m.ep_t.reshape(eg.shape)) < 1e-8   # oracle state: [element][gp][component]
    ctx.close()
