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


NH_MIXED_CFG = {   # config 5's path in small: compressible Neo-Hooke, F33 ramp with P11 = P22 = 0 (test_MixedBCs_LargeStrain.json, load case 1)
    "microstructure": {"L": [1.0, 1.5, 2.0]}, "problem_type": "mechanical", "strain_type": "large", "FE_type": "HEX8", "method": "cg",
    "materials": [{"phases": [0, 1], "matmodel": "CompressibleNeoHookean",
                   "material_properties": {"bulk_modulus": [62.5, 222.222], "shear_modulus": [28.8462, 166.6667]}}],
    "error_parameters": {"measure": "Linfinity", "type": "absolute", "tolerance": 1e-9}, "n_it": 400,
    "macroscale_loading": [{"strain_indices": [1, 2, 3, 5, 6, 7, 8], "stress_indices": [0, 4],
                            "strain": [[0, 0, 0, 0, 0, 0, 1.02], [0, 0, 0, 0, 0, 0, 1.04]], "stress": [[0.0, 0.0], [0.0, 0.0]]}]}


def run_gpu_load_cases(ms, cfg, max_steps=None, on_step=None, make_ctx=None):
    """Python mirror of runSolver (src/main.cpp:9-46) + MixedBCController::activate (mixedBCs.h:180-226) driving the
    CUDA library through its C ABI.  Returns (per load case -> per step dict) like fans_oracle.run_load_cases."""
    problem = cfg["problem_type"]
    strain_type = cfg.get("strain_type", "small")
    n_str = 3 if problem == "thermal" else (9 if strain_type == "large" else 6)
    ep = cfg["error_parameters"]
    ls = cfg.get("linesearch_parameters", {})
    results = []
    for entry in cfg["macroscale_loading"]:
        # the oracle object is used here ONLY as a parameter parser (phase descriptors, kapparef): n_it=0, never solved
        par = fo.OracleSolver(ms, cfg["microstructure"]["L"], problem, cfg["materials"], cfg.get("FE_type", "HEX8"), cfg["method"],
                              strain_type, ep, 0, None, cfg.get("reference_material"))
        ctx = ctx_from_oracle(par) if make_ctx is None else make_ctx(par)   # make_ctx: e.g. one slab of a multi-GPU run
        mbc = None
        if isinstance(entry, dict):
            mbc = fo.MixedBC(entry["strain_indices"], entry["stress_indices"], entry.get("strain", []), entry.get("stress", []), n_str)
            mbc.finalize(par.kapparef)
            n_steps = mbc.n_steps
        else:
            n_steps = len(entry)
        g0_vec = g0_prev = None
        steps = []
        for t in range(n_steps if max_steps is None else min(n_steps, max_steps)):
            if mbc is not None:
                if t == 0:
                    g0_vec = np.zeros(n_str)
                    if n_str == 9:
                        g0_vec[[0, 4, 8]] = 1.0
                    g0_prev = g0_vec.copy()
                else:
                    g0_vec = ctx.get_gradient()
                    delta = g0_vec - g0_prev
                    g0_prev = g0_vec.copy()
                    for k in mbc.idx_F:
                        g0_vec[k] += delta[k]
                for i, k in enumerate(mbc.idx_E):
                    g0_vec[k] = mbc.F_E_path[t, i]
                ctx.set_gradient(g0_vec)
                ctx.set_mixed_bc(mbc.idx_F, mbc.M, mbc.P_F_path[t] if mbc.idx_F else [])
                ctx.update_mixed_bc()
            else:
                ctx.set_mixed_bc(None)
                ctx.set_gradient(np.asarray(entry[t], dtype=np.float64))
            r = ctx.solve(cfg["method"], cfg["n_it"], ep["tolerance"], ep["measure"], ep["type"], ls.get("max_iter", 5), ls.get("tol", 1e-2))
            res = {"iters": r["iters"], "err_all": r["err_all"], "g0": ctx.get_gradient(), "n_residual_evals": r["n_residual_evals"]}
            if on_step is not None:
                on_step(ctx, len(results), t, res)
            else:
                res["stress_average"] = ctx.homogenized_stress()
            steps.append(res)
            if cfg.get("extrapolate_displacement", True):
                ctx.extrapolate_displacement()
        results.append(steps)
        last = ctx
    return results, last
