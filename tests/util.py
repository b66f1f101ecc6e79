"""Test helpers: build a libfans_gpu context that mirrors an OracleSolver (oracle is the CHECKER, never the product)."""
import numpy as np

import fans_oracle as fo
from fans_b200 import _lib as L


def phase_descs_from_oracle(sol):
    """One fans_phase_desc per phase id, derived from the oracle's parsed models (same numbers the reference's
    Matmodel constructors compute)."""
    descs = []
    for p in range(sol.n_phases):
        m = sol.models[sol.phase_model[p]]
        i = sol.phase_local[p]
        d = L.PhaseDesc()
        d.local_mat = i
        d.group_n_mat = m.n_mat
        par = []
        if m.is_linear:
            d.model = L.MAT_LINEAR
            par = list(np.asarray(m.phase_kappa(i), dtype=np.float64).reshape(-1))
        elif isinstance(m, fo.PseudoPlasticLinearHardening):
            d.model = L.MAT_PP_LIN
            par = [m.K[i], m.G[i], m.sy[i], m.H[i], m.eps_crit[i], m.E_s[i]]
        elif isinstance(m, fo.PseudoPlasticNonLinearHardening):
            d.model = L.MAT_PP_NONLIN
            par = [m.K[i], m.G[i], m.sy[i], m.n_exp[i], m.eps0[i], m.eps_crit[i]]
        elif isinstance(m, fo.J2ViscoPlastic_NonLinearIsotropicHardening):
            d.model = L.MAT_J2_NONLIN
            par = [m.K[i], m.G[i], m.sy[i], m.Kiso[i], m.H[i], m.eta[i], m.dt, m.sinf[i], m.delta[i]]
        elif isinstance(m, fo.J2ViscoPlastic_LinearIsotropicHardening):
            d.model = L.MAT_J2_LIN
            par = [m.K[i], m.G[i], m.sy[i], m.Kiso[i], m.H[i], m.eta[i], m.dt]
        elif isinstance(m, fo.J2PlasticityNew_LinearIsotropicHardening):
            d.model = L.MAT_J2NEW
            par = [m.K[i], m.G[i], m.sy[i], m.Kiso[i]]
        elif isinstance(m, fo.SaintVenantKirchhoff):
            d.model = L.MAT_SVK
            par = [m.lam[i], m.mu[i]]
        elif isinstance(m, fo.CompressibleNeoHookean):
            d.model = L.MAT_NEOHOOKE
            par = [m.lam[i], m.mu[i]]
        else:
            raise TypeError(type(m))
        for k, v in enumerate(par):
            d.params[k] = float(v)
        descs.append(d)
    return descs


def ctx_from_oracle(sol):
    ctx = L.Context((sol.nx, sol.ny, sol.nz), sol.L, sol.h, sol.n_str, sol.FE_type)
    ctx.set_microstructure(sol.ms)
    ctx.set_materials(phase_descs_from_oracle(sol))
    ctx.set_reference_stiffness(sol.kapparef)
    ctx.set_gradient(sol.g0)
    return ctx


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.abs(b).max()
    return np.abs(a - b).max() / (den if den > 0 else 1.0)


def two_phase_ms(n, seed=0, shape=None):
    """random blobby two-phase image (not symmetric, so axis mix-ups show)"""
    rng = np.random.default_rng(seed)
    shape = shape or (n, n, n)
    f = rng.standard_normal(shape)
    F = np.fft.fftn(f)
    k = [np.fft.fftfreq(s) * s for s in shape]
    K2 = k[0][:, None, None] ** 2 + k[1][None, :, None] ** 2 + k[2][None, None, :] ** 2
    g = np.fft.ifftn(F * np.exp(-K2 / 8.0)).real
    return (g > np.quantile(g, 0.6)).astype(np.uint16)


THERMAL = [{"phases": [0, 1], "matmodel": "LinearThermalIsotropic", "material_properties": {"conductivity": [1.0, 10.0]}}]
ELASTIC = [{"phases": [0, 1], "matmodel": "LinearElasticIsotropic",
            "material_properties": {"bulk_modulus": [62.5, 222.222], "shear_modulus": [28.8462, 166.6667]}}]
EP = {"measure": "Linfinity", "type": "absolute", "tolerance": 1e-10}
