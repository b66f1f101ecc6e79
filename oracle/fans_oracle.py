"""
CPU ORACLE for the FANS per-iteration solve loop.  *** TEST INFRASTRUCTURE ONLY ***

This file is a NumPy restatement of the reference algorithm (DataAnalyticsEngineering/FANS v0.6.2).
It exists only so that `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference`
legs of `bench.py` can check / time the CUDA product path against it.  Nothing under `fans_b200/`
imports it, and the product path has no CPU fallback.

Pinning status (see tests/test_oracle_kat.py and DESIGN.md section "Oracle"):
  * pinned by the two known answers embedded in the reference's own test inputs
      KAT-1  test/input_files/test_MixedBCs.json:71-75          (pseudo-plastic / elastic, HEX8)
      KAT-2  test/input_files/test_MixedBCs_LargeStrain.json:53-59 (compressible Neo-Hooke, HEX8)
    and by the reference's pytest invariants (test/pytest/*.py) restated in tests/.
  * J2 plasticity, pseudo-plastic plastic branch, BBAR and thermal values: PARITY UNPINNED by any
    number the reference repository holds (it cannot be built here: no MPI/FFTW/HDF5/Eigen).

Every function cites the reference file:line it follows (paths relative to /root/reference).
Memory layout: fields are numpy arrays u[x, y, z, howmany]; ms[x, y, z] (uint16) as in
src/reader.cpp:385-394 (logical X-Y-Z order, z fastest).
"""
from __future__ import annotations

import math
import numpy as np

try:  # same DFT, all host threads (the CPU-baseline leg of bench.py times this path)
    import scipy.fft as _fft
    _FFT_KW = {"workers": -1}
except Exception:  # pragma: no cover
    _fft, _FFT_KW = np.fft, {}

SQRT_HALF = 7.071067811865476e-01  # include/matmodel.h:288


# --------------------------------------------------------------------------------------------
# B matrices                                                         include/matmodel.h:104-188
# --------------------------------------------------------------------------------------------
def compute_basic_B(x, y, z, l_e):
    """Gradient of the 8 trilinear shape functions at (x,y,z) in [0,1]^3. matmodel.h:157-188."""
    lx, ly, lz = l_e
    out = np.zeros((3, 8))
    out[0] = np.array([-(1 - y) * (1 - z), (1 - y) * (1 - z), -y * (1 - z), y * (1 - z),
                       -(1 - y) * z, (1 - y) * z, -y * z, y * z]) / lx
    out[1] = np.array([-(1 - x) * (1 - z), -x * (1 - z), (1 - x) * (1 - z), x * (1 - z),
                       -(1 - x) * z, -x * z, (1 - x) * z, x * z]) / ly
    out[2] = np.array([-(1 - x) * (1 - y), -x * (1 - y), -(1 - x) * y, -x * y,
                       (1 - x) * (1 - y), x * (1 - y), (1 - x) * y, x * y]) / lz
    return out


def compute_B(kind, x, y, z, l_e):
    """kind: 'thermal' (3x8), 'small' (6x24 Mandel, matmodel.h:284-304),
    'large' (9x24 row-major F, LargeStrainMechModel.h:209-225)."""
    b = compute_basic_B(x, y, z, l_e)
    if kind == "thermal":
        return b
    if kind == "small":
        out = np.zeros((6, 24))
        for q in range(8):
            out[0, 3 * q + 0] = b[0, q]
            out[1, 3 * q + 1] = b[1, q]
            out[2, 3 * q + 2] = b[2, q]
            out[3, 3 * q + 0] = SQRT_HALF * b[1, q]
            out[4, 3 * q + 0] = SQRT_HALF * b[2, q]
            out[5, 3 * q + 1] = SQRT_HALF * b[2, q]
            out[3, 3 * q + 1] = SQRT_HALF * b[0, q]
            out[4, 3 * q + 2] = SQRT_HALF * b[0, q]
            out[5, 3 * q + 2] = SQRT_HALF * b[1, q]
        return out
    if kind == "large":
        out = np.zeros((9, 24))
        for i in range(3):
            for J in range(3):
                for node in range(8):
                    out[3 * i + J, 3 * node + i] = b[J, node]
        return out
    raise ValueError(kind)


def construct_B(kind, fe_type, l_e):
    """Returns list B_int[gp] (n_str x 8h). matmodel.h:104-155."""
    n_str = {"thermal": 3, "small": 6, "large": 9}[kind]
    if fe_type in ("HEX8", "BBAR"):
        xi_p = 0.5 + math.sqrt(3.0) / 6.0
        xi_m = 0.5 - math.sqrt(3.0) / 6.0
        xi = [(xi_m, xi_m, xi_m), (xi_p, xi_m, xi_m), (xi_m, xi_p, xi_m), (xi_p, xi_p, xi_m),
              (xi_m, xi_m, xi_p), (xi_p, xi_m, xi_p), (xi_m, xi_p, xi_p), (xi_p, xi_p, xi_p)]
        B_int = []
        if fe_type == "BBAR":
            B_vol = compute_B(kind, 0.5, 0.5, 0.5, l_e)
            for p in range(8):
                B_full = compute_B(kind, *xi[p], l_e)
                if n_str > 3:
                    Bp = np.zeros_like(B_full)
                    vol_full = (B_full[0] + B_full[1] + B_full[2]) / 3.0
                    vol_bar = (B_vol[0] + B_vol[1] + B_vol[2]) / 3.0
                    for r in range(3):
                        Bp[r] = B_full[r] - vol_full + vol_bar
                    Bp[3:] = B_full[3:]
                    B_int.append(Bp)
                else:
                    B_int.append(B_full)
        else:
            for p in range(8):
                B_int.append(compute_B(kind, *xi[p], l_e))
        return B_int
    if fe_type == "HEX8R":
        return [compute_B(kind, 0.5, 0.5, 0.5, l_e)]
    raise RuntimeError("Unknown FE_type: '%s'. Supported types: HEX8, HEX8R, BBAR" % fe_type)


# --------------------------------------------------------------------------------------------
# Material models (vectorised over elements; each follows get_sigma of its reference class)
# eps/sigma arrays are shaped (M, n_gp, n_str); `mat` is the local_mat_id of each element (M,)
# --------------------------------------------------------------------------------------------
class Model:
    kind = None
    n_str = None
    is_linear = False
    has_history = False

    def __init__(self, props, n_mat, B_int, v_e):
        self.props = props
        self.n_mat = n_mat
        self.B_int = B_int
        self.n_gp = len(B_int)
        self.v_e = v_e

    def phase_kappa(self, i):  # material tangent of a linear model (n_str x n_str)
        raise NotImplementedError

    def phase_stiffness(self):
        """LinearModel::phase_stiffness, e.g. LinearElastic.h:27-40 / LinearThermal.h:23-31."""
        out = []
        for i in range(self.n_mat):
            K = np.zeros((self.B_int[0].shape[1],) * 2)
            kap = self.phase_kappa(i)
            for Bp in self.B_int:
                K += Bp.T @ kap @ Bp * self.v_e / self.n_gp
            out.append(K)
        return out

    def init_history(self, n_el):
        pass

    def commit_history(self):
        pass

    def sigma(self, eps, mat, el):
        raise NotImplementedError

    def reference_stiffness(self):
        raise NotImplementedError


def _vec(props, key, n=None):
    v = np.atleast_1d(np.asarray(props[key], dtype=np.float64))
    return v


class LinearThermalIsotropic(Model):
    """include/material_models/LinearThermal.h:7-46"""
    kind, n_str, is_linear = "thermal", 3, True

    def __init__(self, props, n_mat, B_int, v_e):
        super().__init__(props, n_mat, B_int, v_e)
        self.k = _vec(props, "conductivity")
        self.n_mat = len(self.k)

    def phase_kappa(self, i):
        return self.k[i] * np.eye(3)

    def sigma(self, eps, mat, el):
        return self.k[mat][:, None, None] * eps

    def reference_stiffness(self):
        return sum(self.phase_kappa(i) for i in range(self.n_mat)) / self.n_mat


class LinearThermalTriclinic(Model):
    """include/material_models/LinearThermal.h:48-118"""
    kind, n_str, is_linear = "thermal", 3, True

    def __init__(self, props, n_mat, B_int, v_e):
        super().__init__(props, n_mat, B_int, v_e)
        keys = ["K_11", "K_12", "K_13", "K_22", "K_23", "K_33"]
        c = np.array([_vec(props, k) for k in keys])
        self.n_mat = c.shape[1]
        self.K = np.zeros((self.n_mat, 3, 3))
        for i in range(self.n_mat):
            self.K[i] = [[c[0, i], c[1, i], c[2, i]], [c[1, i], c[3, i], c[4, i]], [c[2, i], c[4, i], c[5, i]]]

    def phase_kappa(self, i):
        return self.K[i]

    def sigma(self, eps, mat, el):
        return np.einsum("mij,mgj->mgi", self.K[mat], eps)

    def reference_stiffness(self):
        return self.K.sum(0) / self.n_mat


class LinearElasticIsotropic(Model):
    """include/material_models/LinearElastic.h:7-75"""
    kind, n_str, is_linear = "small", 6, True

    def __init__(self, props, n_mat, B_int, v_e):
        super().__init__(props, n_mat, B_int, v_e)
        self.K = _vec(props, "bulk_modulus")
        self.mu = _vec(props, "shear_modulus")
        self.n_mat = len(self.K)
        self.lam = self.K - (2.0 / 3.0) * self.mu

    def phase_kappa(self, i):
        k = np.zeros((6, 6))
        k[:3, :3] = self.lam[i]
        k += 2 * self.mu[i] * np.eye(6)
        return k

    def sigma(self, eps, mat, el):
        lam = self.lam[mat][:, None]
        buf2 = 2 * self.mu[mat][:, None, None]
        buf1 = lam * (eps[..., 0] + eps[..., 1] + eps[..., 2])
        s = buf2 * eps
        s[..., :3] += buf1[..., None]
        return s

    def reference_stiffness(self):  # LinearElastic.h:55-69 : (max+min)/2
        lam_ref = (self.lam.max() + self.lam.min()) / 2
        mu_ref = (self.mu.max() + self.mu.min()) / 2
        k = np.zeros((6, 6))
        k[:3, :3] = lam_ref
        k += 2 * mu_ref * np.eye(6)
        return k


class LinearElasticTriclinic(Model):
    """include/material_models/LinearElastic.h:77-159"""
    kind, n_str, is_linear = "small", 6, True

    def __init__(self, props, n_mat, B_int, v_e):
        super().__init__(props, n_mat, B_int, v_e)
        keys = ["C_%d%d" % (r + 1, c + 1) for r in range(6) for c in range(r, 6)]
        cc = np.array([_vec(props, k) for k in keys])
        self.n_mat = cc.shape[1]
        self.C = np.zeros((self.n_mat, 6, 6))
        for i in range(self.n_mat):
            k = 0
            for r in range(6):
                for c in range(r, 6):
                    self.C[i, r, c] = cc[k, i]
                    self.C[i, c, r] = cc[k, i]
                    k += 1

    def phase_kappa(self, i):
        return self.C[i]

    def sigma(self, eps, mat, el):
        return np.einsum("mij,mgj->mgi", self.C[mat], eps)

    def reference_stiffness(self):
        return self.C.sum(0) / self.n_mat


def _iso_ref(K, G):
    """PseudoPlastic.h:43-53, J2Plasticity.h:113-123 : arithmetic mean of K and G."""
    Kbar = K.sum() / len(K)
    Gbar = G.sum() / len(G)
    lam = Kbar - 2.0 * Gbar / 3.0
    k = np.zeros((6, 6))
    k[:3, :3] = lam
    k[np.arange(6), np.arange(6)] += 2.0 * Gbar
    return k


class PseudoPlasticBase(Model):
    """include/material_models/PseudoPlastic.h:22-76"""
    kind, n_str = "small", 6

    def __init__(self, props, n_mat, B_int, v_e):
        super().__init__(props, n_mat, B_int, v_e)
        self.K = _vec(props, "bulk_modulus")
        self.G = _vec(props, "shear_modulus")
        self.sy = _vec(props, "yield_stress")
        self.n_mat = len(self.K)
        self.plastic_flag = None

    def init_history(self, n_el):  # PseudoPlastic.h:36-39
        self.plastic_flag = np.zeros((n_el, self.n_gp), dtype=np.int32)

    def reference_stiffness(self):
        return _iso_ref(self.K, self.G)

    def _dev(self, eps):
        treps = eps[..., 0] + eps[..., 1] + eps[..., 2]
        dev = eps.copy()
        dev[..., :3] -= (1.0 / 3.0) * treps[..., None]
        return treps, dev


class PseudoPlasticLinearHardening(PseudoPlasticBase):
    """PseudoPlastic.h:78-124"""

    def __init__(self, props, n_mat, B_int, v_e):
        super().__init__(props, n_mat, B_int, v_e)
        self.H = _vec(props, "hardening_parameter")
        self.eps_crit = math.sqrt(2.0 / 3.0) * self.sy / (2.0 * self.G)
        self.E_s = (3.0 * self.G) / (3.0 * self.G + self.H)
        self.a = 2.0 / 3
        self.b = math.sqrt(self.a)

    def sigma(self, eps, mat, el):
        treps, dev = self._dev(eps)
        nrm = np.sqrt((dev * dev).sum(-1))
        m = mat[:, None]
        buf1 = self.K[m] * treps
        elastic = nrm <= self.eps_crit[m]
        with np.errstate(divide="ignore", invalid="ignore"):
            buf2_pl = (self.b * self.sy[m] + self.a * self.E_s[m] * self.H[m] * (nrm - self.eps_crit[m])) / nrm
        buf2 = np.where(elastic, 2.0 * self.G[m], buf2_pl)
        self.plastic_flag[el] = np.where(elastic, m, self.n_mat + m)
        s = buf2[..., None] * dev
        s[..., :3] += buf1[..., None]
        return s


class PseudoPlasticNonLinearHardening(PseudoPlasticBase):
    """PseudoPlastic.h:126-173"""

    def __init__(self, props, n_mat, B_int, v_e):
        super().__init__(props, n_mat, B_int, v_e)
        self.n_exp = _vec(props, "hardening_exponent")
        self.eps0 = _vec(props, "eps_0")
        self.eps_crit = self.eps0 * np.power(self.sy / (3.0 * self.G * self.eps0), 1.0 / (1.0 - self.n_exp))

    def sigma(self, eps, mat, el):
        treps, dev = self._dev(eps)
        dn = np.sqrt((dev * dev).sum(-1))
        nrm = math.sqrt(2.0 / 3.0) * dn
        m = mat[:, None]
        buf1 = self.K[m] * treps
        elastic = nrm <= self.eps_crit[m]
        with np.errstate(divide="ignore", invalid="ignore"):
            buf2_pl = math.sqrt(2.0 / 3.0) * self.sy[m] * np.power(nrm / self.eps0[m], self.n_exp[m])
            s_pl = buf2_pl[..., None] * dev / dn[..., None]
        s_el = 2.0 * self.G[m][..., None] * dev
        s = np.where(elastic[..., None], s_el, s_pl)
        s[..., :3] += buf1[..., None]
        self.plastic_flag[el] = np.where(elastic, m, self.n_mat + m)
        return s


class J2Plasticity(Model):
    """include/material_models/J2Plasticity.h:7-163 (base), history per GP for EVERY element."""
    kind, n_str, has_history = "small", 6, True

    def __init__(self, props, n_mat, B_int, v_e):
        super().__init__(props, n_mat, B_int, v_e)
        self.K = _vec(props, "bulk_modulus")
        self.G = _vec(props, "shear_modulus")
        self.sy = _vec(props, "yield_stress")
        self.Kiso = _vec(props, "isotropic_hardening_parameter")
        self.H = _vec(props, "kinematic_hardening_parameter")
        self.eta = _vec(props, "viscosity")
        self.dt = float(props["time_step"])
        self.n_mat = len(self.K)
        self.c23 = math.sqrt(2.0 / 3.0)

    def init_history(self, n_el):  # J2Plasticity.h:47-56
        g = self.n_gp
        self.ep = np.zeros((n_el, g, 6)); self.ep_t = np.zeros((n_el, g, 6))
        self.psi = np.zeros((n_el, g)); self.psi_t = np.zeros((n_el, g))
        self.psib = np.zeros((n_el, g, 6)); self.psib_t = np.zeros((n_el, g, 6))

    def commit_history(self):  # J2Plasticity.h:58-63
        self.ep_t = self.ep.copy(); self.psi_t = self.psi.copy(); self.psib_t = self.psib.copy()

    def reference_stiffness(self):
        return _iso_ref(self.K, self.G)

    def q_trial(self, psi, m):
        raise NotImplementedError

    def gamma(self, f_trial, m, psi_t):
        raise NotImplementedError

    def sigma(self, eps, mat, el):  # J2Plasticity.h:65-108
        m = mat[:, None]
        G = self.G[m]
        ee = eps - self.ep_t[el]
        treps = ee[..., 0] + ee[..., 1] + ee[..., 2]
        st = 2 * G[..., None] * ee
        st[..., :3] += ((self.K[m] - 2.0 * G / 3.0) * treps)[..., None]
        dev = st.copy()
        dev[..., :3] -= (st[..., :3].sum(-1) / 3.0)[..., None]
        psi_t = self.psi_t[el]
        q_tr = self.q_trial(psi_t, m)
        qbar = -(2.0 / 3.0) * self.H[m][..., None] * self.psib_t[el]
        dmq = dev - qbar
        nrm = np.sqrt((dmq * dmq).sum(-1))
        with np.errstate(divide="ignore", invalid="ignore"):
            n = np.where((nrm < 1e-12)[..., None], 0.0, dmq / nrm[..., None])
        f_trial = nrm - self.c23 * (self.sy[m] - q_tr)
        gam = np.where(f_trial < 0, 0.0, self.gamma(f_trial, m, psi_t))
        st = st - (gam * 2 * G)[..., None] * n
        self.ep[el] = self.ep_t[el] + gam[..., None] * n
        self.psi[el] += gam * self.c23          # quirk: accumulates on every call (J2Plasticity.h:103)
        self.psib[el] -= gam[..., None] * n     # quirk: accumulates on every call (J2Plasticity.h:104)
        return st


class J2ViscoPlastic_LinearIsotropicHardening(J2Plasticity):
    """J2Plasticity.h:165-178"""

    def q_trial(self, psi, m):
        return -self.Kiso[m] * psi

    def gamma(self, f_trial, m, psi_t):
        return f_trial / (2 * self.G[m] + (2.0 / 3.0) * (self.Kiso[m] + self.H[m]) + self.eta[m] / self.dt)


class J2ViscoPlastic_NonLinearIsotropicHardening(J2Plasticity):
    """J2Plasticity.h:180-243"""

    def __init__(self, props, n_mat, B_int, v_e):
        super().__init__(props, n_mat, B_int, v_e)
        self.sinf = _vec(props, "saturation_stress")
        self.delta = _vec(props, "saturation_exponent")
        self.denom = 2 * self.G + (2.0 / 3.0) * (self.Kiso + self.H) + self.eta / self.dt
        self.sdiff = self.c23 * (self.sinf - self.sy)

    def q_trial(self, psi, m):
        return -self.Kiso[m] * psi - (self.sinf[m] - self.sy[m]) * (1 - np.exp(-self.delta[m] * psi))

    def gamma(self, f_trial, m, psi_t):  # J2Plasticity.h:207-223
        den = np.broadcast_to(self.denom[m], f_trial.shape)
        sd = np.broadcast_to(self.sdiff[m], f_trial.shape)
        dl = np.broadcast_to(self.delta[m], f_trial.shape)
        gam = np.zeros_like(f_trial)
        ginc = np.ones_like(f_trial)
        active = np.ones(f_trial.shape, dtype=bool)
        for _ in range(10):
            active = active & (ginc > 1e-10)
            if not active.any():
                break
            g = f_trial - gam * den - sd * (-np.exp(-dl * (psi_t + self.c23 * gam)) + np.exp(-dl * psi_t))
            # (2 / 3) is integer division == 0 in the reference (J2Plasticity.h:217)
            dg = -den - 0.0
            gi = -g / dg
            ginc = np.where(active, gi, ginc)
            gam = np.where(active, gam + gi, gam)
        return gam


class J2PlasticityNew_LinearIsotropicHardening(Model):
    """include/material_models/J2PlasticityNew.h:7-150"""
    kind, n_str, has_history = "small", 6, True

    def __init__(self, props, n_mat, B_int, v_e):
        super().__init__(props, n_mat, B_int, v_e)
        self.K = _vec(props, "bulk_modulus")
        self.G = _vec(props, "shear_modulus")
        self.sy = _vec(props, "yield_stress")
        self.Kiso = _vec(props, "isotropic_hardening_parameter")
        self.n_mat = len(self.K)
        self.c23 = math.sqrt(2.0 / 3.0)

    def init_history(self, n_el):
        g = self.n_gp
        self.ep = np.zeros((n_el, g, 6)); self.ep_t = np.zeros((n_el, g, 6))
        self.q = np.zeros((n_el, g)); self.q_t = np.zeros((n_el, g))

    def commit_history(self):
        self.ep_t = self.ep.copy(); self.q_t = self.q.copy()

    def reference_stiffness(self):
        return _iso_ref(self.K, self.G)

    def sigma(self, eps, mat, el):
        m = mat[:, None]
        G = self.G[m]
        ep_in = self.ep_t[el]
        s = 2.0 * G[..., None] * (eps - ep_in)
        lam = self.K[m] - 2.0 / 3.0 * G
        s[..., :3] += (lam * eps[..., :3].sum(-1))[..., None]
        sig_t = s.copy()
        sig_t[..., :3] -= (s[..., :3].sum(-1) / 3.0)[..., None]
        s_t = np.sqrt((sig_t * sig_t).sum(-1))
        q_in = self.q_t[el]
        sy = self.sy[m] + q_in * self.Kiso[m]
        dsy = self.Kiso[m]
        phi = s_t - self.c23 * sy
        dgam = np.maximum(0.0, phi / (2.0 * G + 2.0 / 3.0 * dsy))
        big = s_t > 1e-12
        with np.errstate(divide="ignore", invalid="ignore"):
            nn = np.where(big[..., None], sig_t / s_t[..., None], 0.0)
        self.ep[el] = ep_in + dgam[..., None] * nn
        self.q[el] = q_in + self.c23 * dgam
        return s - (dgam * 2.0 * G)[..., None] * nn


MANDEL_IJ = [(0, 0), (1, 1), (2, 2), (0, 1), (0, 2), (1, 2)]


def compute_spatial_tangent(F, S, C_mandel):
    """LargeStrainMechModel.h:105-180 verbatim, including the P<=Q-only sum."""
    A = np.zeros((9, 9))

    def mandel(a, b):
        for idx, (p, q) in enumerate(MANDEL_IJ):
            if (p == a and q == b) or (p == b and q == a):
                return idx
        return -1

    for i in range(3):
        for J in range(3):
            row = 3 * i + J
            for k in range(3):
                for L in range(3):
                    col = 3 * k + L
                    if i == k:
                        A[row, col] += S[L, J]
                    for M in range(3):
                        MJ = mandel(M, J)
                        for P in range(3):
                            for Q in range(P, 3):
                                PQ = mandel(P, Q)
                                C_val = C_mandel[MJ, PQ]
                                if MJ >= 3:
                                    C_val /= math.sqrt(2.0)
                                if PQ >= 3:
                                    C_val /= math.sqrt(2.0)
                                dE = 0.0
                                if Q == L:
                                    dE += 0.5 * F[k, P]
                                if P == L:
                                    dE += 0.5 * F[k, Q]
                                A[row, col] += F[i, M] * C_val * dE
    return A


class LargeStrainModel(Model):
    """include/LargeStrainMechModel.h:11-207 : eps == F (row-major 9), sigma == P = F S."""
    kind, n_str = "large", 9

    def __init__(self, props, n_mat, B_int, v_e):
        super().__init__(props, n_mat, B_int, v_e)
        self.Kb = _vec(props, "bulk_modulus")
        self.mu = _vec(props, "shear_modulus")
        self.n_mat = len(self.Kb)
        self.lam = self.Kb - (2.0 / 3.0) * self.mu

    def compute_S(self, F, m):
        raise NotImplementedError

    def material_tangent_at_identity(self, i):
        raise NotImplementedError

    def sigma(self, eps, mat, el):
        F = eps.reshape(eps.shape[:-1] + (3, 3))
        S = self.compute_S(F, mat[:, None])
        P = F @ S
        return P.reshape(eps.shape)

    def reference_stiffness(self):
        k = np.zeros((9, 9))
        I = np.eye(3)
        for i in range(self.n_mat):
            S = self.compute_S(I[None, None], np.array([[i]]))[0, 0]
            k += compute_spatial_tangent(I, S, self.material_tangent_at_identity(i))
        return k / self.n_mat


class SaintVenantKirchhoff(LargeStrainModel):
    """include/material_models/SaintVenantKirchhoff.h:33-66"""

    def compute_S(self, F, m):
        C = np.swapaxes(F, -1, -2) @ F
        E = 0.5 * (C - np.eye(3))
        trE = np.trace(E, axis1=-2, axis2=-1)
        return (self.lam[m] * trE)[..., None, None] * np.eye(3) + 2.0 * self.mu[m][..., None, None] * E

    def material_tangent_at_identity(self, i):
        C = np.zeros((6, 6))
        C[:3, :3] = self.lam[i]
        return C + 2.0 * self.mu[i] * np.eye(6)


class CompressibleNeoHookean(LargeStrainModel):
    """include/material_models/CompressibleNeoHookean.h:35-108"""

    def compute_S(self, F, m):
        C = np.swapaxes(F, -1, -2) @ F
        J = np.linalg.det(F)
        if np.any(J <= 0.0):
            raise RuntimeError("Negative Jacobian determinant in CompressibleNeoHookean!")
        logJ = np.log(J)
        Cinv = np.linalg.inv(C)
        return (self.lam[m] * logJ)[..., None, None] * Cinv + self.mu[m][..., None, None] * (np.eye(3) - Cinv)

    def material_tangent_at_identity(self, i):  # CompressibleNeoHookean.h:50-92 at F = I
        Cinv = np.eye(3)
        logJ = 0.0
        fac = [1.0, 1.0, 1.0, math.sqrt(2.0), math.sqrt(2.0), math.sqrt(2.0)]
        cm = np.array([Cinv[0, 0], Cinv[1, 1], Cinv[2, 2], fac[3] * Cinv[0, 1], fac[4] * Cinv[0, 2], fac[5] * Cinv[1, 2]])
        PP1 = np.outer(cm, cm)
        PP2 = np.zeros((6, 6))
        for a in range(6):
            I_, J_ = MANDEL_IJ[a]
            for b in range(6):
                K_, L_ = MANDEL_IJ[b]
                PP2[a, b] = (Cinv[I_, K_] * Cinv[J_, L_] + Cinv[I_, L_] * Cinv[J_, K_]) * fac[a] * fac[b]
        return self.lam[i] * PP1 + (self.mu[i] - self.lam[i] * logJ) * PP2


MODEL_REGISTRY = {  # include/setup.h:21-73
    "thermal": {"LinearThermalIsotropic": LinearThermalIsotropic, "LinearThermalTriclinic": LinearThermalTriclinic},
    "small": {"LinearElasticIsotropic": LinearElasticIsotropic, "LinearElasticTriclinic": LinearElasticTriclinic,
              "PseudoPlasticLinearHardening": PseudoPlasticLinearHardening,
              "PseudoPlasticNonLinearHardening": PseudoPlasticNonLinearHardening,
              "J2ViscoPlastic_LinearIsotropicHardening": J2ViscoPlastic_LinearIsotropicHardening,
              "J2ViscoPlastic_NonLinearIsotropicHardening": J2ViscoPlastic_NonLinearIsotropicHardening,
              "J2PlasticityNew_LinearIsotropicHardening": J2PlasticityNew_LinearIsotropicHardening},
    "large": {"SaintVenantKirchhoff": SaintVenantKirchhoff, "CompressibleNeoHookean": CompressibleNeoHookean},
}


# --------------------------------------------------------------------------------------------
# Mixed boundary conditions                                          include/mixedBCs.h:15-226
# --------------------------------------------------------------------------------------------
class MixedBC:
    def __init__(self, strain_indices, stress_indices, strain, stress, n_str):
        self.idx_E = list(strain_indices)
        self.idx_F = list(stress_indices)
        present = [0] * n_str
        for k in self.idx_E:
            if k < 0 or k >= n_str:
                raise RuntimeError("strain index out of range")
            present[k] = 1
        for k in self.idx_F:
            if k < 0 or k >= n_str:
                raise RuntimeError("stress index out of range")
            if present[k]:
                raise RuntimeError("index appears in both strain_indices and stress_indices")
            present[k] = 1
        if not all(present):
            raise RuntimeError("each component must be either strain- or stress-controlled")
        n_steps = 0
        if self.idx_E:
            n_steps = len(strain)
        if self.idx_F:
            n_steps = max(n_steps, len(stress))
        if n_steps == 0:
            raise RuntimeError("mixed BC: at least one of strain/stress must have timesteps")
        self.F_E_path = np.zeros((n_steps, len(self.idx_E)))
        self.P_F_path = np.zeros((n_steps, len(self.idx_F)))
        for t in range(n_steps):
            if self.idx_E:
                self.F_E_path[t] = strain[t]
            if self.idx_F:
                self.P_F_path[t] = stress[t]
        # NOTE mixedBCs.h / reader.cpp:161 : lc.n_steps = F_E_path.rows()
        self.n_steps = self.F_E_path.shape[0]

    def finalize(self, C0):  # mixedBCs.h:30-46
        n_str = C0.shape[0]
        self.Q_E = np.zeros((n_str, len(self.idx_E)))
        for c, k in enumerate(self.idx_E):
            self.Q_E[k, c] = 1.0
        self.Q_F = np.zeros((n_str, len(self.idx_F)))
        for c, k in enumerate(self.idx_F):
            self.Q_F[k, c] = 1.0
        if self.idx_F:
            self.M = np.linalg.pinv(self.Q_F.T @ C0 @ self.Q_F)
        else:
            self.M = np.zeros((0, 0))


# --------------------------------------------------------------------------------------------
# Solver                                       include/solver.h, solverCG.h, solverFP.h
# --------------------------------------------------------------------------------------------
NODE_SHIFT = [((i >> 0) & 1, (i >> 1) & 1, (i >> 2) & 1) for i in range(8)]  # solver.h:333-340


class OracleSolver:
    """Restates Solver<howmany,n_str> + SolverCG/SolverFP + MaterialManager + MixedBCController.

    Parameters mirror the reference's JSON keys (src/reader.cpp:63-183)."""

    def __init__(self, ms, L, problem_type, materials, FE_type="HEX8", method="cg", strain_type="small",
                 error_parameters=None, n_it=100, linesearch_parameters=None, reference_material=None,
                 n_ranks=1, verbose=False):
        self.ms = np.ascontiguousarray(ms).astype(np.uint16)
        self.nx, self.ny, self.nz = self.ms.shape
        self.N = self.nx * self.ny * self.nz
        self.L = list(L)
        self.l_e = [self.L[0] / self.nx, self.L[1] / self.ny, self.L[2] / self.nz]
        self.v_e = self.l_e[0] * self.l_e[1] * self.l_e[2]
        if problem_type == "thermal":
            self.kind, self.h, self.n_str = "thermal", 1, 3
        elif problem_type == "mechanical" and strain_type == "small":
            self.kind, self.h, self.n_str = "small", 3, 6
        elif problem_type == "mechanical" and strain_type == "large":
            self.kind, self.h, self.n_str = "large", 3, 9
        else:
            raise ValueError(problem_type + " is not a valid problem type")
        self.FE_type = FE_type
        self.method = method
        ep = error_parameters or {"measure": "Linfinity", "type": "absolute", "tolerance": 1e-10}
        self.measure, self.err_type, self.TOL = ep["measure"], ep["type"], float(ep["tolerance"])
        self.n_it = int(n_it)
        ls = linesearch_parameters or {}
        self.ls_max_iter = int(ls.get("max_iter", 5))
        self.ls_tol = float(ls.get("tol", 1e-2))
        self.n_ranks = n_ranks
        self.verbose = verbose

        self.B_int = construct_B(self.kind, FE_type, self.l_e)
        self.n_gp = len(self.B_int)
        self.B = np.vstack(self.B_int)  # (n_str*n_gp) x 8h   matmodel.h:141,146,151

        # ---- MaterialManager (MaterialManager.h:38-159) ----
        max_phase = max(max(g["phases"]) for g in materials)
        self.n_phases = max_phase + 1
        self.models = []
        self.phase_model = [-1] * self.n_phases
        self.phase_local = [-1] * self.n_phases
        for gi, g in enumerate(materials):
            cls = MODEL_REGISTRY[self.kind].get(g["matmodel"])
            if cls is None:
                raise ValueError(g["matmodel"] + " is not a valid matmodel")
            mdl = cls(g["material_properties"], len(g["phases"]), self.B_int, self.v_e)
            self.models.append(mdl)
            for i, p in enumerate(g["phases"]):
                if p < 0 or p >= self.n_phases or self.phase_model[p] >= 0:
                    raise RuntimeError("MaterialManager: Invalid or duplicate phase %d" % p)
                self.phase_model[p] = gi
                self.phase_local[p] = i
        if any(p < 0 for p in self.phase_model):
            raise RuntimeError("MaterialManager: Phase not assigned")
        self.all_linear = all(m.is_linear for m in self.models)
        if reference_material is not None:  # MaterialManager.h:179-196
            self.kapparef = np.array(reference_material, dtype=np.float64)
            np.linalg.cholesky(self.kapparef)
        else:
            self.kapparef = sum(m.reference_stiffness() for m in self.models) / len(self.models)

        msf = self.ms.reshape(-1).astype(np.int64)
        pm = np.array(self.phase_model)[msf]
        self.el_of_model = [np.nonzero(pm == gi)[0] for gi in range(len(self.models))]
        self.local_of_el = np.array(self.phase_local)[msf]
        for m in self.models:
            m.init_history(self.N)  # solver.h:139 : for EVERY element
        self.phase_K = {}
        for gi, m in enumerate(self.models):
            if m.is_linear:
                self.phase_K[gi] = np.array(m.phase_stiffness())

        h = self.h
        self.u = np.zeros((self.nx, self.ny, self.nz, h))
        self.u_prev = np.zeros_like(self.u)
        self.r = np.zeros_like(self.u)
        self.g0 = np.zeros(self.n_str)
        self.err_all = np.zeros(self.n_it + 1)
        self.iter = 0
        self.n_residual_evals = 0
        self.mixed_active = False
        self.gamma_hat = self.compute_fundamental_solution()

    # ---------------- MaterialManager::set_gradient (MaterialManager.h:216-221) ----------
    def set_gradient(self, g0):
        self.g0 = np.asarray(g0, dtype=np.float64).copy()

    # ---------------- element gather / scatter (solver.h:313-385) ------------------------
    def _gather(self, u):
        cols = []
        for (a, b, c) in NODE_SHIFT:
            cols.append(np.roll(u, shift=(-a, -b, -c), axis=(0, 1, 2)))
        return np.stack(cols, axis=3).reshape(self.N, 8 * self.h)  # node-major, component-minor

    def _scatter(self, res_e):
        res = res_e.reshape(self.nx, self.ny, self.nz, 8, self.h)
        r = np.zeros((self.nx, self.ny, self.nz, self.h))
        for i, (a, b, c) in enumerate(NODE_SHIFT):
            r += np.roll(res[:, :, :, i, :], shift=(a, b, c), axis=(0, 1, 2))
        return r

    def _sigma_all(self, eps):
        """eps (N, n_gp, n_str) -> sigma, dispatching per model like MaterialManager::get_info."""
        sig = np.zeros_like(eps)
        for gi, m in enumerate(self.models):
            el = self.el_of_model[gi]
            if len(el):
                sig[el] = m.sigma(eps[el], self.local_of_el[el], el)
        return sig

    # ---------------- compute_residual (solver.h:229-280, matmodel.h:190-201) -------------
    def compute_residual(self, u):
        self.n_residual_evals += 1
        ue = self._gather(u)
        ue = ue - np.tile(ue[:, :self.h], (1, 8))           # solver.h:252  u_i - u_0
        eps = ue @ self.B.T + np.tile(self.g0, self.n_gp)   # matmodel.h:194
        sig = self._sigma_all(eps.reshape(self.N, self.n_gp, self.n_str))
        res_e = sig.reshape(self.N, -1) @ self.B * (self.v_e / self.n_gp)  # matmodel.h:199
        return self._scatter(res_e)

    # ---------------- linear operator (solverCG.h:98-103) --------------------------------
    def apply_linear(self, d):
        ue = self._gather(d)
        ue = ue - np.tile(ue[:, :self.h], (1, 8))
        res_e = np.zeros_like(ue)
        for gi in self.phase_K:
            el = self.el_of_model[gi]
            loc = self.local_of_el[el]
            for m in np.unique(loc):          # one dense product per phase: res_e = K_phase ue  (LinearModel::phase_stiffness)
                sel = el[loc == m]
                res_e[sel] = ue[sel] @ self.phase_K[gi][m].T
        return self._scatter(res_e)

    # ---------------- Green operator (solver.h:144-204) -----------------------------------
    def reference_element_stiffness(self):
        """matmodel.h:237-253 (component-major reordering)."""
        n = 8 * self.h
        tmp = np.zeros((n, n))
        for Bp in self.B_int:
            tmp += Bp.T @ self.kapparef @ Bp * self.v_e / self.n_gp
        out = np.zeros((n, n))
        h = self.h
        for i in range(n):
            for j in range(n):
                out[(i % h) * 8 + i // h, (j % h) * 8 + j // h] = tmp[i, j]
        return out

    def compute_fundamental_solution(self):
        """Returns Gamma_hat[kx, ky, kz, h, h] (already divided by N, solver.h:198);
        pinv with ABSOLUTE singular value cut 1e-14 (solver.h:89-96,189-191); xi = 0 left zero (:169)."""
        h = self.h
        Ker0 = self.reference_element_stiffness()
        nzc = self.nz // 2 + 1
        ex = np.exp(2j * np.pi * np.arange(self.nx) / self.nx)[:, None, None]
        ey = np.exp(2j * np.pi * np.arange(self.ny) / self.ny)[None, :, None]
        ez = np.exp(2j * np.pi * np.arange(nzc) / self.nz)[None, None, :]
        one = np.ones((self.nx, self.ny, nzc), dtype=complex)
        A = np.stack([one, ex * one, ey * one, ex * ey * one, ez * one, ex * ez * one, ez * ey * one, ex * ey * ez], axis=-1)
        AA = A.real[..., :, None] * A.real[..., None, :] + A.imag[..., :, None] * A.imag[..., None, :]
        block = np.zeros((self.nx, self.ny, nzc, h, h))
        for i in range(h):
            for j in range(i, h):
                block[..., i, j] = (Ker0[8 * i:8 * i + 8, 8 * j:8 * j + 8] * AA).sum((-1, -2))
                block[..., j, i] = block[..., i, j]
        U, s, Vt = np.linalg.svd(block)
        with np.errstate(divide="ignore"):
            sinv = np.where(s > 1e-14, 1.0 / s, 0.0)
        G = np.einsum("...ji,...j,...kj->...ik", Vt, sinv, U)  # V diag(sinv) U^T
        G[0, 0, 0] = 0.0
        self.n_singular = int((s <= 1e-14).any(-1).sum())
        self.gamma_sv = s  # singular values per frequency (tests use them to spot blocks sitting ON the 1e-14 cut)
        return G / float(self.N)

    def convolution(self, r):
        """solver.h:387-412 : unnormalised r2c, per-frequency Gamma_hat multiply, unnormalised c2r."""
        rhat = _fft.rfftn(r, axes=(0, 1, 2), **_FFT_KW)
        count = self.ny * self.nx * (self.nz // 2 + 1)
        shat = np.einsum("xyzij,xyzj->xyzi", self.gamma_hat, rhat)
        if count % 2 == 1:
            # solver.h:400 integer division drops the last frequency (layout [ky][kx][kz]) -> left as r_hat
            shat[-1, -1, -1] = rhat[-1, -1, -1]
        return _fft.irfftn(shat, s=(self.nx, self.ny, self.nz), axes=(0, 1, 2), **_FFT_KW) * float(self.N)

    # ---------------- compute_error (solver.h:414-452) -------------------------------------
    def compute_error(self, r):
        slabs = np.array_split(r, self.n_ranks, axis=0)  # MAX over ranks for every measure (solver.h:430)
        if self.measure == "L1":
            err = max(np.abs(s).sum() for s in slabs)
        elif self.measure == "L2":
            err = max(math.sqrt((s * s).sum()) for s in slabs)
        elif self.measure == "Linfinity":
            err = max(np.abs(s).max() for s in slabs)
        else:
            raise RuntimeError("Unknown measure type: " + self.measure)
        self.err_all[self.iter] = err
        err0 = self.err_all[0]
        with np.errstate(divide="ignore", invalid="ignore"):
            err_rel = 100.0 if self.iter == 0 else err / err0
        if self.verbose:
            print("it %3d .... err %16.8e" % (self.iter, err))
        if self.err_type == "absolute":
            return err
        if self.err_type == "relative":
            return err_rel
        raise RuntimeError("Unknown error type: " + self.err_type)

    # ---------------- solve (solver.h:282-300) ------------------------------------------------
    def solve(self):
        self.err_all = np.zeros(self.n_it + 1)
        if self.method == "cg":
            self._solve_cg()
        elif self.method == "fp":
            self._solve_fp()
        else:
            raise ValueError(self.method + " is not a valid method")
        for m in self.models:  # MaterialManager::update_internal_variables
            m.commit_history()

    @staticmethod
    def _dot(a, b):
        return float((a * b).sum())

    def _solve_cg(self):  # solverCG.h:61-117
        self.alpha_warm = 0.1
        s = np.zeros_like(self.u)
        d = np.zeros_like(self.u)
        self.r = self.compute_residual(self.u)
        self.iter = 0
        err_rel = self.compute_error(self.r)
        delta = 1.0
        while self.iter < self.n_it and err_rel > self.TOL:
            deltamid = self._dot(self.r, s)
            s = -self.convolution(self.r)
            delta0 = delta
            delta = self._dot(self.r, s)
            with np.errstate(divide="ignore", invalid="ignore"):
                beta = np.fmax(0.0, np.float64(delta - deltamid) / np.float64(delta0))
            d = s + beta * d
            if self.all_linear and not self.mixed_active:
                rnew = self.apply_linear(d)
                with np.errstate(divide="ignore", invalid="ignore"):
                    alpha = np.float64(delta) / np.float64(self._dot(d, rnew))
                self.r = self.r - alpha * rnew
                self.u = self.u - alpha * d
            else:
                self._line_search_secant(d)
            self.iter += 1
            err_rel = self.compute_error(self.r)
        self.d_last = d

    def _line_search_secant(self, d):  # solverCG.h:119-160
        err = 10.0
        it = 0
        alpha_prev = 0.0
        alpha_curr = self.alpha_warm
        rpd = self._dot(self.r, d)
        self.u = self.u + d * alpha_curr
        self.update_mixed_bc()
        rnew = self.compute_residual(self.u)
        r1pd = self._dot(rnew, d)
        while it < self.ls_max_iter and err > self.ls_tol:
            denom = r1pd - rpd
            if abs(denom) < 1e-14 * (abs(r1pd) + abs(rpd)):
                break
            alpha_next = alpha_curr - r1pd * (alpha_curr - alpha_prev) / denom
            if alpha_next <= 0.0:
                alpha_next = 0.5 * (alpha_prev + alpha_curr)
            err = abs(alpha_next - alpha_curr)
            self.u = self.u + d * (alpha_next - alpha_curr)
            alpha_prev = alpha_curr
            rpd = r1pd
            alpha_curr = alpha_next
            it += 1
            self.update_mixed_bc()
            rnew = self.compute_residual(self.u)
            r1pd = self._dot(rnew, d)
        self.alpha_warm = 0.1 if (it == self.ls_max_iter and err > self.ls_tol) else alpha_curr
        self.r = rnew

    def _solve_fp(self):  # solverFP.h:32-55
        self.r = self.compute_residual(self.u)
        self.iter = 0
        err_rel = self.compute_error(self.r)
        while self.iter < self.n_it and err_rel > self.TOL:
            self.r = self.convolution(self.r)
            self.u = self.u - self.r
            self.update_mixed_bc()
            self.r = self.compute_residual(self.u)
            self.iter += 1
            err_rel = self.compute_error(self.r)

    # ---------------- extrapolateDisplacement (solver.h:302-311) --------------------------------
    def extrapolate_displacement(self):
        delta = self.u - self.u_prev
        self.u_prev = self.u.copy()
        self.u = self.u + delta

    # ---------------- strain / stress sweep (solver.h:707-737, matmodel.h:202-225) ---------------
    def strain_stress(self):
        """Per-element GP-averaged strain and stress with ABSOLUTE ue (solver.h:507,723).
        Also returns the GP arrays.  Calls get_sigma => history side effects as in the reference."""
        ue = self._gather(self.u)
        eps = (ue @ self.B.T + np.tile(self.g0, self.n_gp)).reshape(self.N, self.n_gp, self.n_str)
        sig = self._sigma_all(eps)
        return eps.sum(1) / self.n_gp, sig.sum(1) / self.n_gp, eps, sig

    def get_homogenized_stress(self):
        _, stress, _, _ = self.strain_stress()
        return stress.sum(0) / self.N

    def get_homogenized_tangent(self, pert_param=1e-6):  # solver.h:739-778
        n = self.n_str
        C = np.zeros((n, n))
        unperturbed = self.get_homogenized_stress()
        g0 = self.g0.copy()
        self.err_type = "relative"            # quirk 7: permanent
        self.TOL = max(1e-6, self.TOL)
        for m in self.models:
            if isinstance(m, J2Plasticity):
                raise RuntimeError("Homogenized tangent computation not implemented for J2Plasticity models.")
        for i in range(n):
            if self.all_linear:
                pert = np.zeros(n)
                pert[i] = 1.0
            else:
                pert = g0.copy()
                pert[i] += pert_param
            self.set_gradient(pert)
            self.mixed_active = False
            self.solve()
            perturbed = self.get_homogenized_stress()
            C[:, i] = perturbed if self.all_linear else (perturbed - unperturbed) / pert_param
        return 0.5 * (C + C.T)

    # ---------------- MixedBCController (mixedBCs.h:150-226) --------------------------------------
    def enable_mixed_bc(self, mbc: MixedBC, t: int):
        self.mixed_active = True
        self.mbc = mbc
        self.step_idx = t
        mbc.finalize(self.kapparef)
        n = self.n_str
        if t == 0:
            self.g0_vec = np.zeros(n)
            if n == 9:
                self.g0_vec[[0, 4, 8]] = 1.0
            self.g0_vec_prev = self.g0_vec.copy()
        else:
            delta = self.g0_vec - self.g0_vec_prev
            self.g0_vec_prev = self.g0_vec.copy()
            for k in mbc.idx_F:
                self.g0_vec[k] += delta[k]
        if mbc.idx_E:
            for i, k in enumerate(mbc.idx_E):
                self.g0_vec[k] = mbc.F_E_path[t, i]
        self.set_gradient(self.g0_vec)
        self.update_mixed_bc()

    def disable_mixed_bc(self):
        self.mixed_active = False

    def update_mixed_bc(self):
        if not self.mixed_active:
            return
        Pbar = self.get_homogenized_stress()
        mbc = self.mbc
        if mbc.idx_F:
            PF = mbc.P_F_path[self.step_idx]
            rhs = PF - mbc.Q_F.T @ Pbar
            self.g0_vec = self.g0_vec + mbc.Q_F @ (mbc.M @ rhs)
        self.set_gradient(self.g0_vec)

    # ---------------- postprocess data sources (solver.h:454-705) ----------------------------------
    def postprocess(self):
        strain, stress, eps_gp, sig_gp = self.strain_stress()
        out = {
            "strain": strain.reshape(self.nx, self.ny, self.nz, self.n_str),
            "stress": stress.reshape(self.nx, self.ny, self.nz, self.n_str),
            "strain_gp": eps_gp.reshape(self.nx, self.ny, self.nz, self.n_gp, self.n_str),
            "stress_gp": sig_gp.reshape(self.nx, self.ny, self.nz, self.n_gp, self.n_str),
            "stress_average": stress.sum(0) / self.N,
            "strain_average": strain.sum(0) / self.N,
            "absolute_error": self.err_all[: self.iter + 1].copy(),
            "displacement_fluctuation": self.u.copy(),
            "residual": self.r.copy(),
        }
        msf = self.ms.reshape(-1)
        n_mat = int(msf.max()) - int(msf.min()) + 1  # reader.cpp:36
        for p in range(n_mat):
            sel = msf == p
            cnt = int(sel.sum())
            out["phase_stress_average_phase%d" % p] = stress[sel].sum(0) / cnt if cnt else np.zeros(self.n_str)
            out["phase_strain_average_phase%d" % p] = strain[sel].sum(0) / cnt if cnt else np.zeros(self.n_str)
        # u_total = g0.X + u~   (solver.h:587-650)
        sa = out["strain_average"]
        X = (np.arange(self.nx) * self.l_e[0] - self.L[0] / 2.0)[:, None, None]
        Y = (np.arange(self.ny) * self.l_e[1] - self.L[1] / 2.0)[None, :, None]
        Z = (np.arange(self.nz) * self.l_e[2] - self.L[2] / 2.0)[None, None, :]
        ut = self.u.copy()
        rs2 = 0.7071067811865475
        if self.h == 3 and self.n_str == 6:
            g11, g22, g33, g12, g13, g23 = sa[0], sa[1], sa[2], sa[3] * rs2, sa[4] * rs2, sa[5] * rs2
            ut[..., 0] += g11 * X + g12 * Y + g13 * Z
            ut[..., 1] += g12 * X + g22 * Y + g23 * Z
            ut[..., 2] += g13 * X + g23 * Y + g33 * Z
        elif self.h == 3:
            ut[..., 0] += (sa[0] - 1.0) * X + sa[1] * Y + sa[2] * Z
            ut[..., 1] += sa[3] * X + (sa[4] - 1.0) * Y + sa[5] * Z
            ut[..., 2] += sa[6] * X + sa[7] * Y + (sa[8] - 1.0) * Z
        else:
            ut[..., 0] += sa[0] * X + sa[1] * Y + sa[2] * Z
        out["displacement"] = ut
        for m in self.models:
            if isinstance(m, PseudoPlasticBase):  # PseudoPlastic.h:55-63
                out["plastic_flag"] = m.plastic_flag.astype(np.float32).mean(1).reshape(self.nx, self.ny, self.nz)
            if isinstance(m, J2Plasticity):  # J2Plasticity.h:245-322 (the *_t copies, GP mean)
                out["plastic_strain"] = m.ep_t.mean(1).reshape(self.nx, self.ny, self.nz, 6)
                out["isotropic_hardening_variable"] = m.psi_t.mean(1).reshape(self.nx, self.ny, self.nz)
                out["kinematic_hardening_variable"] = m.psib_t.mean(1).reshape(self.nx, self.ny, self.nz, 6)
            if isinstance(m, J2PlasticityNew_LinearIsotropicHardening):
                out["plastic_strain"] = m.ep_t.mean(1).reshape(self.nx, self.ny, self.nz, 6)
                out["isotropic_hardening_variable"] = m.q_t.mean(1).reshape(self.nx, self.ny, self.nz)
        return out


# --------------------------------------------------------------------------------------------
# Driver mirroring runSolver (src/main.cpp:9-46)
# --------------------------------------------------------------------------------------------
def sphere_microstructure(n=32, radius_frac=0.4):
    """Formula of test/microstructures/sphere32.h5 (/sphere/32x32x32/ms): phase 1 where
    (i-c)^2+(j-c)^2+(k-c)^2 <= (0.4 n)^2, c=(n-1)/2. For n=32: 8744 voxels (checked in tests)."""
    c = (n - 1) / 2.0
    i = np.arange(n) - c
    d2 = i[:, None, None] ** 2 + i[None, :, None] ** 2 + i[None, None, :] ** 2
    return (d2 <= (radius_frac * n) ** 2).astype(np.uint16)


def run_load_cases(ms, cfg, n_ranks=1, max_steps=None, verbose=False, on_step=None):
    """cfg: dict with the reference's JSON keys. Returns list (per load case) of list (per step) of dicts."""
    problem = cfg["problem_type"]
    strain_type = cfg.get("strain_type", "small")
    n_str = 3 if problem == "thermal" else (9 if strain_type == "large" else 6)
    results = []
    for lc_idx, entry in enumerate(cfg["macroscale_loading"]):
        sol = OracleSolver(ms, cfg["microstructure"]["L"], problem, cfg["materials"], cfg.get("FE_type", "HEX8"),
                           cfg["method"], strain_type, cfg["error_parameters"], cfg["n_it"],
                           cfg.get("linesearch_parameters"), cfg.get("reference_material"), n_ranks, verbose)
        if isinstance(entry, dict):
            mbc = MixedBC(entry["strain_indices"], entry["stress_indices"], entry.get("strain", []),
                          entry.get("stress", []), n_str)
            n_steps = mbc.n_steps
        else:
            mbc = None
            n_steps = len(entry)
        steps = []
        for t in range(n_steps if max_steps is None else min(n_steps, max_steps)):
            if mbc is not None:
                sol.enable_mixed_bc(mbc, t)
            else:
                sol.set_gradient(entry[t])
            sol.solve()
            res = {"iters": sol.iter, "err_all": sol.err_all[: sol.iter + 1].copy(), "g0": sol.g0.copy(),
                   "n_residual_evals": sol.n_residual_evals}
            if on_step is not None:
                on_step(sol, lc_idx, t, res)
            else:
                res["stress_average"] = sol.get_homogenized_stress()
            steps.append(res)
            if cfg.get("extrapolate_displacement", True):
                sol.extrapolate_displacement()
        results.append(steps)
        last_solver = sol
    return results, last_solver
