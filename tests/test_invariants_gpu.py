"""Oracle-INDEPENDENT checks of the GPU material laws and element formulations: properties the reference's equations imply, evaluated
on the library's own Gauss-point outputs (strain_gp / stress_gp / history *_gp).  They pin the laws for which the reference repository
holds no known-answer value (J2 plasticity, plastic branch of the pseudo-plastic laws, B-bar, thermal) without going through
oracle/fans_oracle.py at all — fans_oracle is imported here only as the parser of material parameters (ctx_from_oracle)."""
import numpy as np
import pytest

import fans_oracle as fo
import util
from fans_b200 import simple

pytestmark = pytest.mark.gpu

K0, G0, SY = 62.5, 28.8462, 0.1
LOAD = [0.002, -0.0005, -0.0004, 0.0006, 0.0, 0.0003]   # ||dev|| = 2.1e-3
S23 = np.sqrt(2.0 / 3.0)


def dev(t):
    """deviator of Mandel vectors (..., 6)"""
    d = t.copy()
    d[..., :3] -= t[..., :3].mean(-1, keepdims=True)
    return d


def smooth_u(shape, amp, seed=0):
    rng = np.random.default_rng(seed)
    n = [np.arange(s) / s for s in shape]
    u = np.zeros(shape + (3,))
    for c in range(3):
        ph = rng.uniform(0, 2 * np.pi, 3)
        u[..., c] = amp * np.sin(2 * np.pi * n[0] + ph[0])[:, None, None] * np.cos(2 * np.pi * n[1] + ph[1])[None, :, None] \
            * np.sin(4 * np.pi * n[2] + ph[2])[None, None, :]
    return u


def one_phase_ctx(mat, shape=(8, 8, 16), fe="HEX8", L=(1.0, 1.0, 1.0)):
    ms = np.zeros(shape, dtype=np.uint16)
    par = fo.OracleSolver(ms, list(L), "mechanical", [mat], fe, "cg", "small", util.EP, 0)   # parameter parsing only
    return util.ctx_from_oracle(par)


@pytest.mark.parametrize("kind", ["j2_lin", "j2_nonlin"])
def test_j2_return_map_invariants(kind):
    """J2Plasticity.h:65-108 after ONE evaluation from a virgin state: plastic incompressibility, psi = sqrt(2/3) gamma,
    psi_bar = -eps_p, Hooke's law on the elastic strain, and the visco-plastic consistency condition
    ||dev sigma - qbar|| - sqrt(2/3) (sigma_y - q) = eta gamma / dt with the UPDATED hardening variables; elastic points obey Hooke."""
    Kiso, Hk, eta, dt = (3.0, 2.0, 1.0, 0.01) if kind == "j2_lin" else (1.5, 2.0, 1.0, 0.01)
    props = {"bulk_modulus": [K0], "shear_modulus": [G0], "yield_stress": [SY], "isotropic_hardening_parameter": [Kiso],
             "kinematic_hardening_parameter": [Hk], "viscosity": [eta], "time_step": dt}
    name = "J2ViscoPlastic_LinearIsotropicHardening"
    sinf, delta = 0.15, 50.0
    if kind == "j2_nonlin":
        name = "J2ViscoPlastic_NonLinearIsotropicHardening"
        props.update(saturation_stress=[sinf], saturation_exponent=[delta])
    ctx = one_phase_ctx({"phases": [0], "matmodel": name, "material_properties": props})
    ctx.set_gradient(list(0.67 * np.array(LOAD)))   # ||dev eps|| ~ 1.4e-3 = the elastic limit sqrt(2/3) sigma_y / (2 G): both branches occur
    ctx.upload("u", smooth_u(ctx.dims, 2e-4))
    _, _, eps, sig = ctx.strain_stress_gp()      # one law call per Gauss point, history written
    ctx.commit_history()
    ep = ctx.get_field("plastic_strain_gp")
    psi = ctx.get_field("isotropic_hardening_variable_gp")
    pb = ctx.get_field("kinematic_hardening_variable_gp")
    gam = np.linalg.norm(ep, axis=-1)
    plastic = gam > 0
    assert 0.05 < plastic.mean() < 0.999, plastic.mean()     # both branches present
    assert np.abs(ep[..., :3].sum(-1)).max() < 1e-16 + 1e-13 * gam.max()          # tr eps_p = 0
    assert np.abs(psi - S23 * gam).max() < 1e-14 * max(gam.max(), 1e-30) + 1e-18    # psi = sqrt(2/3) gamma
    assert np.abs(pb + ep).max() < 1e-14 * gam.max()                                 # psi_bar = -gamma n = -eps_p
    # Hooke on the elastic strain: sigma = K tr(eps) 1 + 2 G dev(eps - eps_p)
    ee = eps - ep
    hooke = 2.0 * G0 * dev(ee)
    hooke[..., :3] += K0 * ee[..., :3].sum(-1, keepdims=True)
    assert np.abs(sig - hooke).max() < 1e-12 * np.abs(sig).max()
    # consistency with the updated variables: q = -K psi [- (s_inf - s_y)(1 - exp(-delta psi))], qbar = -(2/3) H psi_bar
    q = -Kiso * psi
    if kind == "j2_nonlin":
        q = q - (sinf - SY) * (1.0 - np.exp(-delta * psi))
    f = np.linalg.norm(dev(sig) + (2.0 / 3.0) * Hk * pb, axis=-1) - S23 * (SY - q)
    scale = S23 * SY
    if kind == "j2_lin":
        assert np.abs(f[plastic] - eta * gam[plastic] / dt).max() < 1e-12 * scale
    else:
        # The reference's Newton loop (J2Plasticity.h:207-223) runs on the SIGNED increment with the slope -den (its "(2 / 3)" is
        # integer 0): the first step overshoots, the second comes back and, being negative, ends the loop.  The consistency condition
        # therefore holds to the square of the contraction factor c = sqrt(2/3) (s_inf - s_y) delta sqrt(2/3) / den, not to round-off ...
        den = 2.0 * G0 + (2.0 / 3.0) * (Kiso + Hk) + eta / dt
        c = (2.0 / 3.0) * (sinf - SY) * delta / den
        f_trial = 2.0 * G0 * np.linalg.norm(dev(eps), axis=-1) - S23 * SY      # virgin state: q = qbar = 0
        assert np.abs(f[plastic] - eta * gam[plastic] / dt).max() < 2.0 * c * c * f_trial[plastic].max()
        # ... and gamma is exactly that two-step iterate
        sd = S23 * (sinf - SY)
        g1 = f_trial / den
        g2 = g1 + (f_trial - g1 * den - sd * (1.0 - np.exp(-delta * S23 * g1))) / den
        assert np.abs(gam[plastic] - g2[plastic]).max() < 1e-12 * gam.max()
    assert f[~plastic].max() <= 1e-14                                               # elastic points are inside the yield surface
    ctx.close()


def test_j2new_return_map_invariants():
    """J2PlasticityNew.h:43-109 (rate-independent, linear isotropic hardening): after the return the stress sits ON the yield surface
    ||dev sigma|| = sqrt(2/3) (sigma_y + K q), q = sqrt(2/3) gamma, tr eps_p = 0."""
    Kiso = 4.0
    ctx = one_phase_ctx({"phases": [0], "matmodel": "J2PlasticityNew_LinearIsotropicHardening",
                         "material_properties": {"bulk_modulus": [K0], "shear_modulus": [G0], "yield_stress": [SY],
                                                 "isotropic_hardening_parameter": [Kiso]}})
    ctx.set_gradient(list(0.67 * np.array(LOAD)))
    ctx.upload("u", smooth_u(ctx.dims, 2e-4, 1))
    _, _, eps, sig = ctx.strain_stress_gp()
    ctx.commit_history()
    ep = ctx.get_field("plastic_strain_gp")
    q = ctx.get_field("isotropic_hardening_variable_gp")
    gam = np.linalg.norm(ep, axis=-1)
    plastic = gam > 0
    assert 0.05 < plastic.mean() < 0.999
    assert np.abs(ep[..., :3].sum(-1)).max() < 1e-13 * gam.max()
    assert np.abs(q - S23 * gam).max() < 1e-14 * gam.max()
    f = np.linalg.norm(dev(sig), axis=-1) - S23 * (SY + Kiso * q)
    assert np.abs(f[plastic]).max() < 1e-12 * SY and f[~plastic].max() <= 1e-14
    ctx.close()


@pytest.mark.parametrize("kind", ["lin", "nonlin"])
def test_pseudoplastic_hardening_curve(kind):
    """PseudoPlastic.h:95-116, 143-168: sigma = K tr(eps) 1 + beta dev(eps) with ||dev sigma|| on the hardening curve, continuous at the
    elastic limit eps_crit = sqrt(2/3) sigma_y / (2 G) (linear) / where the power law meets 2 G ||e|| (the model's own eps_crit)."""
    H = 5.0
    if kind == "lin":
        mat = {"phases": [0], "matmodel": "PseudoPlasticLinearHardening",
               "material_properties": {"bulk_modulus": [K0], "shear_modulus": [G0], "yield_stress": [SY], "hardening_parameter": [H]}}
    else:
        mat = {"phases": [0], "matmodel": "PseudoPlasticNonLinearHardening",
               "material_properties": {"bulk_modulus": [K0], "shear_modulus": [G0], "yield_stress": [SY], "hardening_exponent": [0.2], "eps_0": [0.01]}}
    ctx = one_phase_ctx(mat)
    sc = 0.67 if kind == "lin" else 0.35     # the power law leaves the elastic branch at eps_eq = 6.7e-4 already
    ctx.set_gradient(list(sc * np.array(LOAD)))
    ctx.upload("u", smooth_u(ctx.dims, 2e-4 * sc, 2))
    _, _, eps, sig = ctx.strain_stress_gp()
    e, s = dev(eps), dev(sig)
    ne, nsig = np.linalg.norm(e, axis=-1), np.linalg.norm(s, axis=-1)
    assert np.abs(sig[..., :3].mean(-1) - K0 * eps[..., :3].sum(-1)).max() < 1e-13 * np.abs(sig).max()   # pressure = K tr(eps)
    assert np.abs(s * ne[..., None] - e * nsig[..., None]).max() < 1e-13 * np.abs(s).max() * ne.max()    # dev sigma || dev eps
    if kind == "lin":
        ecrit = S23 * SY / (2.0 * G0)
        Es = 3.0 * G0 / (3.0 * G0 + H)
        curve = np.where(ne <= ecrit, 2.0 * G0 * ne, S23 * SY + (2.0 / 3.0) * Es * H * (ne - ecrit))
        assert abs(2.0 * G0 * ecrit - S23 * SY) < 1e-15                  # continuity at the elastic limit
        plastic = ne > ecrit
    else:
        ecrit = 0.01 * (SY / (3.0 * G0 * 0.01)) ** (1.0 / (1.0 - 0.2))    # PseudoPlastic.h:138-140, in terms of eps_eq = sqrt(2/3) ||e||
        eeq = S23 * ne
        plastic = eeq > ecrit
        curve = np.where(plastic, S23 * SY * (eeq / 0.01) ** 0.2, 2.0 * G0 * ne)
        assert abs(2.0 * G0 * ecrit / S23 - S23 * SY * (ecrit / 0.01) ** 0.2) < 1e-14     # the two branches meet at eps_crit
    assert 0.05 < plastic.mean() < 0.999
    assert np.abs(nsig - curve).max() < 1e-12 * nsig.max()
    ctx.close()


def test_bbar_constant_dilatation():
    """B-bar (matmodel.h:113-140): the volumetric strain of every Gauss point of an element is the element-centre value, the
    deviatoric strain is that of the plain HEX8 element."""
    mat = {"phases": [0], "matmodel": "LinearElasticIsotropic", "material_properties": {"bulk_modulus": [K0], "shear_modulus": [G0]}}
    u = smooth_u((8, 8, 16), 1e-3, 3)
    out = {}
    for fe in ("BBAR", "HEX8", "HEX8R"):
        ctx = one_phase_ctx(mat, fe=fe, L=(1.0, 1.5, 2.0))
        ctx.set_gradient([0.001, -0.002, 0.003, 0.0015, -0.0025, 0.001])
        ctx.upload("u", u)
        out[fe] = ctx.strain_stress_gp()[2]
        ctx.close()
    tr = out["BBAR"][..., :3].sum(-1)                        # [x][y][z][gp]
    assert np.abs(tr - tr[..., :1]).max() < 1e-15 + 1e-13 * np.abs(tr).max()
    assert np.abs(tr[..., 0] - out["HEX8R"][..., 0, :3].sum(-1)).max() < 1e-13 * np.abs(tr).max()   # = the centre (reduced) value
    assert np.abs(dev(out["BBAR"]) - dev(out["HEX8"])).max() < 1e-13 * np.abs(out["HEX8"]).max()


@pytest.mark.parametrize("fe", ["HEX8", "HEX8R", "BBAR"])
def test_thermal_laminate_is_exact(fe):
    """Layers normal to x: the effective conductivity is the harmonic mean across and the arithmetic mean along the layers, and the
    voxel-conforming trilinear solution is exact — q_bar = K_eff g_bar to solver tolerance (LinearThermal.h:35-40)."""
    if fe == "BBAR":
        pytest.skip("B-bar only changes mechanical elements (matmodel.h:113)")
    k = np.array([1.0, 10.0])
    ms = np.zeros((16, 8, 8), dtype=np.uint16)
    ms[5:11] = 1
    f1 = ms.mean()
    ctx = simple.linear_thermal_context(ms, [1.0, 1.0, 1.0], k, fe)
    g = np.array([0.01, 0.02, -0.01])
    ctx.set_gradient(g)
    res = ctx.solve("cg", 200, 1e-12, "Linfinity", "absolute")
    q = ctx.homogenized_stress()
    k_series = 1.0 / ((1 - f1) / k[0] + f1 / k[1])
    k_par = (1 - f1) * k[0] + f1 * k[1]
    assert res["iters"] < 200
    assert np.abs(q - np.array([k_series, k_par, k_par]) * g).max() < 1e-9 * np.abs(q).max()
    ctx.close()


def test_elastic_laminate_uniaxial_is_exact():
    """Mechanical counterpart: layers normal to x under eps_xx only — the stress sigma_xx is uniform and equals the series (Reuss)
    combination of the constrained moduli (lambda + 2 mu): pins Gamma, the element stiffness and the CG path against a closed form."""
    Kb, Gs = np.array([62.5, 222.222]), np.array([28.8462, 166.6667])
    M = Kb + 4.0 / 3.0 * Gs
    ms = np.zeros((16, 8, 8), dtype=np.uint16)
    ms[3:9] = 1
    f1 = ms.mean()
    ctx = simple.linear_elastic_context(ms, [1.0, 1.0, 1.0], Kb, Gs, "HEX8")
    ctx.set_gradient([0.001, 0, 0, 0, 0, 0])
    ctx.solve("cg", 300, 1e-12, "Linfinity", "absolute")
    sig = ctx.homogenized_stress()
    m_eff = 1.0 / ((1 - f1) / M[0] + f1 / M[1])
    assert abs(sig[0] - m_eff * 0.001) < 1e-9 * abs(sig[0])
    _, stress = ctx.strain_stress()
    assert np.abs(stress[..., 0] - sig[0]).max() < 1e-8 * abs(sig[0])     # sigma_xx is uniform
    ctx.close()
