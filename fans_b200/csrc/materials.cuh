// materials.cuh — Gauss-point material laws evaluated in registers (one call = one Matmodel::get_sigma).
// Each branch cites the reference get_sigma it restates.  History lives in COMPACT SoA arrays over the elements whose phase
// carries history (hidx[element] = compact index, built in api.cu; elastic phases such as fibres cost no history memory):
//   hist[(var*ngp + gp)*nh + hel]   (current)  and  hist_t (committed),  compact element index fastest => coalesced.
#pragma once
#include "common.cuh"


#define SQRT_HALF 7.071067811865476e-01          // include/matmodel.h:288
#define SQRT_TWO_THIRDS 0.816496580927726        // sqrt(2/3)

// strain vector from the displacement gradient Hm[c][j] = d u_c / d x_j
//   thermal: grad T (matmodel.h:157-188) ; small strain: Mandel (matmodel.h:284-304) ; large: row-major (LargeStrainMechModel.h:209-225)
template <int H, int NSTR>
__device__ __forceinline__ void strain_from_grad(const double (&Hm)[H][3], double (&e)[NSTR])
{
    if (NSTR == 3) {
        e[0] = Hm[0][0];
        e[1] = Hm[0][1];
        e[2] = Hm[0][2];
    } else if (NSTR == 6) {
        e[0] = Hm[0][0];
        e[1] = Hm[H > 1 ? 1 : 0][1];
        e[2] = Hm[H > 2 ? 2 : 0][2];
        e[3 % NSTR] = SQRT_HALF * (Hm[0][1] + Hm[H > 1 ? 1 : 0][0]);
        e[4 % NSTR] = SQRT_HALF * (Hm[0][2] + Hm[H > 2 ? 2 : 0][0]);
        e[5 % NSTR] = SQRT_HALF * (Hm[H > 1 ? 1 : 0][2] + Hm[H > 2 ? 2 : 0][1]);
    } else {
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int J = 0; J < 3; ++J) e[(3 * i + J) % NSTR] = Hm[i % H][J];
    }
}

// T[c][j] such that (B^T sigma)[h*a + c] = sum_j dN_a/dx_j T[c][j]
template <int H, int NSTR>
__device__ __forceinline__ void stress_tensor(const double (&s)[NSTR], double (&T)[H][3])
{
    if (NSTR == 3) {
        T[0][0] = s[0];
        T[0][1] = s[1];
        T[0][2] = s[2];
    } else if (NSTR == 6) {
        T[0][0] = s[0];
        T[H > 1 ? 1 : 0][1] = s[1];
        T[H > 2 ? 2 : 0][2] = s[2];
        const double s01 = SQRT_HALF * s[3 % NSTR], s02 = SQRT_HALF * s[4 % NSTR], s12 = SQRT_HALF * s[5 % NSTR];
        T[0][1] = s01;
        T[H > 1 ? 1 : 0][0] = s01;
        T[0][2] = s02;
        T[H > 2 ? 2 : 0][0] = s02;
        T[H > 1 ? 1 : 0][2] = s12;
        T[H > 2 ? 2 : 0][1] = s12;
    } else {
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int J = 0; J < 3; ++J) T[i % H][J] = s[(3 * i + J) % NSTR];
    }
}

__device__ __forceinline__ void inv_sym3(const double C[6] /*00,11,22,01,02,12*/, double det, double Ci[6])
{
    const double id = 1.0 / det;
    Ci[0] = (C[1] * C[2] - C[5] * C[5]) * id;
    Ci[1] = (C[0] * C[2] - C[4] * C[4]) * id;
    Ci[2] = (C[0] * C[1] - C[3] * C[3]) * id;
    Ci[3] = (C[4] * C[5] - C[3] * C[2]) * id;
    Ci[4] = (C[3] * C[5] - C[4] * C[1]) * id;
    Ci[5] = (C[3] * C[4] - C[0] * C[5]) * id;
}

// History values of ONE Gauss point staged ahead of time in thread-private shared-memory slots (sweep.cu issues the cp.async of
// Gauss point g+1 while g is evaluated): slot v < 13 = committed value of variable v (hist_t).  s == nullptr: read global memory
// directly.  The CURRENT values of psi / psi_bar (which J2Plasticity accumulates on every call) are never read: the increment goes
// out as a fire-and-forget reduction, and only when the Gauss point yields.
#define FANS_HIST_STAGE_SLOTS 13
struct HistStage {
    const double *s;
    int stride;
    __device__ __forceinline__ double operator()(int slot) const { return s[slot * stride]; }
};

// One Gauss point.  e: strain-like input, s: stress-like output.  LAW >= 0: the model is known at compile time (the caller has
// dispatched on pd.model once per element, so the Gauss-point loop carries no model branches); LAW = -1: dispatch here.
#define FANS_LAW_IS(m) ((LAW < 0) ? (pd.model == (m)) : (LAW == (m)))
template <int NSTR, int LAW = -1>
__device__ __forceinline__ void material_law(const PhaseDev &pd, const double (&e)[NSTR], double (&s)[NSTR], double *hist,
                                             const double *hist_t, int *pflag, size_t nloc, size_t nh, int ngp, int gp, size_t el,
                                             size_t hel, bool write_state, int *fault, const HistStage hs = HistStage{nullptr, 0})
{
    // nloc / el address the dense per-element arrays (plastic_flag), nh / hel the compact history arrays
    const double *P = pd.params;
    if (FANS_LAW_IS(FANS_MAT_LINEAR)) {
        if (pd.lin_iso) {
            if (NSTR == 6) {
                // LinearElasticIsotropic::get_sigma (LinearElastic.h:43-53): params = lambda, 2 mu
                const double buf1 = P[0] * (e[0] + e[1] + e[2]), buf2 = P[1];
#pragma unroll
                for (int i = 0; i < NSTR; ++i) s[i] = (i < 3 ? buf1 : 0.0) + buf2 * e[i];
                return;
            }
            if (NSTR == 3) {
                // LinearThermalIsotropic::get_sigma (LinearThermal.h:35-40): params = conductivity
#pragma unroll
                for (int i = 0; i < NSTR; ++i) s[i] = P[0] * e[i];
                return;
            }
        }
        // sigma = C eps with the phase tangent (LinearThermal.h:35-40,104-107; LinearElastic.h:43-53,141-144)
        const double *C = pd.tangent;
#pragma unroll
        for (int i = 0; i < NSTR; ++i) {
            double a = 0.0;
#pragma unroll
            for (int j = 0; j < NSTR; ++j) a = fma(__ldg(&C[i * NSTR + j]), e[j], a);
            s[i] = a;
        }
        return;
    }
    if (NSTR == 6) {
        if (FANS_LAW_IS(FANS_MAT_PSEUDOPLASTIC_LINEAR) || FANS_LAW_IS(FANS_MAT_PSEUDOPLASTIC_NONLIN)) {
            // PseudoPlastic.h:95-116 (linear hardening), :143-168 (power law)
            const double K = P[0], G = P[1], sy = P[2];
            const double treps = e[0] + e[1] + e[2];
            double dev[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) dev[i % NSTR] = e[i % NSTR];
            dev[0] -= (1.0 / 3.0) * treps;
            dev[1] -= (1.0 / 3.0) * treps;
            dev[2] -= (1.0 / 3.0) * treps;
            double n2 = 0.0;
#pragma unroll
            for (int i = 0; i < 6; ++i) n2 += dev[i] * dev[i];
            const double dn = sqrt(n2);
            const double buf1 = K * treps;
            double buf2;
            bool elastic;
            if (FANS_LAW_IS(FANS_MAT_PSEUDOPLASTIC_LINEAR)) {
                const double Hh = P[3], eps_crit = P[4], E_s = P[5];
                elastic = dn <= eps_crit;
                buf2 = elastic ? 2.0 * G : (SQRT_TWO_THIRDS * sy + (2.0 / 3.0) * E_s * Hh * (dn - eps_crit)) / dn;
            } else {
                const double nexp = P[3], eps0 = P[4], eps_crit = P[5];
                const double nrm = SQRT_TWO_THIRDS * dn;
                elastic = nrm <= eps_crit;
                buf2 = elastic ? 2.0 * G : SQRT_TWO_THIRDS * sy * pow(nrm / eps0, nexp) / dn;
            }
#pragma unroll
            for (int i = 0; i < 6; ++i) s[i % NSTR] = buf2 * dev[i];
            s[0] += buf1;
            s[1] += buf1;
            s[2] += buf1;
            if (write_state && pflag) pflag[(size_t)gp * nloc + el] = elastic ? pd.local_mat : pd.group_n_mat + pd.local_mat;
            return;
        }
        if (FANS_LAW_IS(FANS_MAT_J2_LINEAR_ISO) || FANS_LAW_IS(FANS_MAT_J2_NONLIN_ISO)) {
            // J2Plasticity.h:65-108; history: plasticStrain(0..5), psi(6), psi_bar(7..12)
            const double K = P[0], G = P[1], sy = P[2], Kiso = P[3], Hk = P[4], eta = P[5], dt = P[6];
            double ept[6], pbt[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                ept[i] = hs.s ? hs(i) : hist_t[((size_t)i * ngp + gp) * nh + hel];
                pbt[i] = hs.s ? hs(7 + i) : hist_t[((size_t)(7 + i) * ngp + gp) * nh + hel];
            }
            const double psit = hs.s ? hs(6) : hist_t[((size_t)6 * ngp + gp) * nh + hel];
            double ee[6], st[6], dev[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) ee[i] = e[i % NSTR] - ept[i];
            const double treps = ee[0] + ee[1] + ee[2];
            const double lam_tr = (K - 2.0 * G / 3.0) * treps;
#pragma unroll
            for (int i = 0; i < 6; ++i) st[i] = (i < 3 ? lam_tr : 0.0) + 2.0 * G * ee[i];
            const double mean = (st[0] + st[1] + st[2]) / 3.0;
#pragma unroll
            for (int i = 0; i < 6; ++i) dev[i] = st[i] - (i < 3 ? mean : 0.0);
            double q_tr;
            if (FANS_LAW_IS(FANS_MAT_J2_LINEAR_ISO)) q_tr = -Kiso * psit;
            else q_tr = -Kiso * psit - (P[7] - sy) * (1.0 - exp(-P[8] * psit));
            double dmq[6], n2 = 0.0;
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                dmq[i] = dev[i] - (-(2.0 / 3.0) * Hk * pbt[i]);
                n2 += dmq[i] * dmq[i];
            }
            const double nrm = sqrt(n2);
            double nvec[6];
            const double inrm = (nrm < 1e-12) ? 0.0 : 1.0 / nrm;  // one reciprocal instead of six FP64 divisions (differs from x / nrm by <= 1 ulp)
#pragma unroll
            for (int i = 0; i < 6; ++i) nvec[i] = dmq[i] * inrm;
            const double f_trial = nrm - SQRT_TWO_THIRDS * (sy - q_tr);
            double gam = 0.0;
            if (!(f_trial < 0)) {
                const double den = 2 * G + (2.0 / 3.0) * (Kiso + Hk) + eta / dt;
                if (FANS_LAW_IS(FANS_MAT_J2_LINEAR_ISO)) {
                    gam = f_trial / den;
                } else {
                    // J2Plasticity.h:207-223: Newton loop on the SIGNED increment; "(2 / 3)" is integer 0 => dg = -den
                    const double sdiff = SQRT_TWO_THIRDS * (P[7] - sy), dl = P[8];
                    double ginc = 1.0;
                    int it = 0;
                    const double e0 = exp(-dl * psit);
                    while (ginc > 1e-10 && it < 10) {
                        const double g = f_trial - gam * den - sdiff * (-exp(-dl * (psit + SQRT_TWO_THIRDS * gam)) + e0);
                        ginc = -g / (-den);
                        gam += ginc;
                        ++it;
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 6; ++i) s[i % NSTR] = st[i] - gam * 2 * G * nvec[i];
            if (write_state) {
                double *hb = hist + (size_t)gp * nh + hel;   // variable v of this Gauss point: hb[v * vs]
                const size_t vs = (size_t)ngp * nh;
#pragma unroll
                for (int i = 0; i < 6; ++i) hb[i * vs] = ept[i] + gam * nvec[i];
                // quirk (J2Plasticity.h:103-104): psi += sqrt(2/3) gamma, psi_bar -= gamma n ACCUMULATE on every call (every residual
                // evaluation of every iteration).  One thread owns the Gauss point, so a reduction (RED.ADD.F64, no load, nothing to
                // wait for) is the same single addition; an elastic point (gamma = 0) adds nothing and touches nothing.
                if (gam != 0.0) {
#pragma unroll
                    for (int i = 0; i < 6; ++i) atomicAdd(&hb[(7 + i) * vs], -(gam * nvec[i]));
                    atomicAdd(&hb[6 * vs], gam * SQRT_TWO_THIRDS);
                }
            }
            return;
        }
        if (FANS_LAW_IS(FANS_MAT_J2NEW_LINEAR_ISO)) {
            // J2PlasticityNew.h:43-109; history: plasticStrain(0..5), q(6)
            const double K = P[0], G = P[1], sy0 = P[2], Kiso = P[3];
            double ep[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) ep[i] = hs.s ? hs(i) : hist_t[((size_t)i * ngp + gp) * nh + hel];
            const double q_in = hs.s ? hs(6) : hist_t[((size_t)6 * ngp + gp) * nh + hel];
            const double lam = K - 2.0 / 3.0 * G;
            const double tr = e[0] + e[1] + e[2];
            double sg[6], sd[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) sg[i] = 2.0 * G * (e[i % NSTR] - ep[i]) + (i < 3 ? lam * tr : 0.0);
            const double mean = (sg[0] + sg[1] + sg[2]) / 3.0;
            double n2 = 0.0;
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                sd[i] = sg[i] - (i < 3 ? mean : 0.0);
                n2 += sd[i] * sd[i];
            }
            const double s_t = sqrt(n2);
            const double syc = sy0 + q_in * Kiso;
            const double phi = s_t - SQRT_TWO_THIRDS * syc;
            const double dgam = fmax(0.0, phi / (2.0 * G + 2.0 / 3.0 * Kiso));
            const double is_t = (s_t > 1e-12) ? 1.0 / s_t : 0.0;
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                const double nn = sd[i] * is_t;
                s[i % NSTR] = sg[i] - dgam * 2.0 * G * nn;
                if (write_state) hist[((size_t)i * ngp + gp) * nh + hel] = ep[i] + dgam * nn;
            }
            if (write_state) hist[((size_t)6 * ngp + gp) * nh + hel] = q_in + SQRT_TWO_THIRDS * dgam;
            return;
        }
    }
    if (NSTR == 9) {
        // LargeStrainMechModel.h:200-206: eps == F (row-major), sigma == P = F S
        const double lam = P[0], mu = P[1];
        double F[3][3];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int J = 0; J < 3; ++J) F[i][J] = e[(3 * i + J) % NSTR];
        double C[6];  // C = F^T F : 00,11,22,01,02,12
        C[0] = F[0][0] * F[0][0] + F[1][0] * F[1][0] + F[2][0] * F[2][0];
        C[1] = F[0][1] * F[0][1] + F[1][1] * F[1][1] + F[2][1] * F[2][1];
        C[2] = F[0][2] * F[0][2] + F[1][2] * F[1][2] + F[2][2] * F[2][2];
        C[3] = F[0][0] * F[0][1] + F[1][0] * F[1][1] + F[2][0] * F[2][1];
        C[4] = F[0][0] * F[0][2] + F[1][0] * F[1][2] + F[2][0] * F[2][2];
        C[5] = F[0][1] * F[0][2] + F[1][1] * F[1][2] + F[2][1] * F[2][2];
        double S[6];
        if (FANS_LAW_IS(FANS_MAT_SVK)) {
            // SaintVenantKirchhoff.h:33-38 : S = lambda tr(E) I + 2 mu E
            const double E0 = 0.5 * (C[0] - 1.0), E1 = 0.5 * (C[1] - 1.0), E2 = 0.5 * (C[2] - 1.0);
            const double ltr = lam * (E0 + E1 + E2);
            S[0] = ltr + 2.0 * mu * E0;
            S[1] = ltr + 2.0 * mu * E1;
            S[2] = ltr + 2.0 * mu * E2;
            S[3] = mu * C[3];
            S[4] = mu * C[4];
            S[5] = mu * C[5];
        } else {
            // CompressibleNeoHookean.h:35-48 : S = lambda ln J C^-1 + mu (I - C^-1); J <= 0 is a fault
            const double J = F[0][0] * (F[1][1] * F[2][2] - F[1][2] * F[2][1]) - F[0][1] * (F[1][0] * F[2][2] - F[1][2] * F[2][0]) +
                             F[0][2] * (F[1][0] * F[2][1] - F[1][1] * F[2][0]);
            if (!(J > 0.0)) {
                if (fault) *fault = FANS_ERR_NEG_JACOBIAN;
            }
            const double logJ = log(J);
            double Ci[6];
            const double detC = C[0] * (C[1] * C[2] - C[5] * C[5]) - C[3] * (C[3] * C[2] - C[5] * C[4]) + C[4] * (C[3] * C[5] - C[1] * C[4]);
            inv_sym3(C, detC, Ci);
            const double a = lam * logJ - mu;
            S[0] = a * Ci[0] + mu;
            S[1] = a * Ci[1] + mu;
            S[2] = a * Ci[2] + mu;
            S[3] = a * Ci[3];
            S[4] = a * Ci[4];
            S[5] = a * Ci[5];
        }
        const double Sm[3][3] = {{S[0], S[3], S[4]}, {S[3], S[1], S[5]}, {S[4], S[5], S[2]}};
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int J = 0; J < 3; ++J) s[(3 * i + J) % NSTR] = F[i][0] * Sm[0][J] + F[i][1] * Sm[1][J] + F[i][2] * Sm[2][J];
        return;
    }
#pragma unroll
    for (int i = 0; i < NSTR; ++i) s[i] = 0.0;
}
