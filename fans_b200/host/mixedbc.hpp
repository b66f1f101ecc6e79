// mixedbc.hpp — mixed stress/strain macroscopic control, mirror of include/mixedBCs.h.
//   MixedBC::from_json   mixedBCs.h:51-133      MixedBC::finalize   mixedBCs.h:30-46
//   LoadCase             mixedBCs.h:139-144
// The per-evaluation update (MixedBCController::update, mixedBCs.h:160-178) runs inside libfans_gpu
// (fans_update_mixed_bc) because it sits inside the line search; activate() lives in solver.hpp.
#pragma once
#include <cmath>
#include <stdexcept>
#include <vector>

#include "json.hpp"
#include "matmodel.hpp"

namespace fans {

// Moore-Penrose pseudo-inverse of a small dense matrix by one-sided Jacobi SVD (stands in for Eigen's
// completeOrthogonalDecomposition().pseudoInverse(), mixedBCs.h:43). Rank cut: eps * max(n) * sigma_max.
static inline Mat pinv_small(const Mat &Ain)
{
    const int m = Ain.r, n = Ain.c;
    if (m == 0 || n == 0) return Mat(n, m);
    Mat U = Ain;  // columns get orthogonalised: A V = U S
    Mat V = Mat::identity(n);
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0;
        for (int p = 0; p < n; ++p)
            for (int q = p + 1; q < n; ++q) {
                double a = 0, b = 0, c = 0;
                for (int i = 0; i < m; ++i) {
                    a += U(i, p) * U(i, p);
                    b += U(i, q) * U(i, q);
                    c += U(i, p) * U(i, q);
                }
                if (std::fabs(c) <= 1e-300 || std::fabs(c) <= 1e-17 * std::sqrt(a * b)) continue;
                off = std::max(off, std::fabs(c) / std::sqrt(a * b));
                const double zeta = (b - a) / (2.0 * c);
                const double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
                const double cs = 1.0 / std::sqrt(1.0 + t * t), sn = cs * t;
                for (int i = 0; i < m; ++i) {
                    const double up = U(i, p), uq = U(i, q);
                    U(i, p) = cs * up - sn * uq;
                    U(i, q) = sn * up + cs * uq;
                }
                for (int i = 0; i < n; ++i) {
                    const double vp = V(i, p), vq = V(i, q);
                    V(i, p) = cs * vp - sn * vq;
                    V(i, q) = sn * vp + cs * vq;
                }
            }
        if (off < 1e-15) break;
    }
    std::vector<double> sig(n);
    double smax = 0.0;
    for (int j = 0; j < n; ++j) {
        double s = 0;
        for (int i = 0; i < m; ++i) s += U(i, j) * U(i, j);
        sig[j] = std::sqrt(s);
        smax = std::max(smax, sig[j]);
    }
    const double tol = 2.220446049250313e-16 * std::max(m, n) * smax;
    Mat P(n, m);  // V S^-1 U^T, with U(:,j) normalised by sig[j]
    for (int j = 0; j < n; ++j) {
        if (sig[j] <= tol) continue;
        const double inv2 = 1.0 / (sig[j] * sig[j]);
        for (int r = 0; r < n; ++r)
            for (int c = 0; c < m; ++c) P(r, c) += V(r, j) * U(c, j) * inv2;
    }
    return P;
}

struct MixedBC {
    std::vector<int> idx_E, idx_F;
    std::vector<std::vector<double>> F_E_path, P_F_path;  // (#steps x |E|), (#steps x |F|)
    size_t n_rows = 0;                                     // F_E_path.rows()
    Mat M;                                                 // (Q_F^T C0 Q_F)^+

    void finalize(const Mat &C0)  // mixedBCs.h:30-46
    {
        const int nF = (int)idx_F.size();
        if (nF > 0) {
            Mat A(nF, nF);
            for (int i = 0; i < nF; ++i)
                for (int j = 0; j < nF; ++j) A(i, j) = C0(idx_F[i], idx_F[j]);
            M = pinv_small(A);
        } else {
            M = Mat(0, 0);
        }
    }

    static MixedBC from_json(const Json &jc, int n_str)  // mixedBCs.h:51-133
    {
        MixedBC bc;
        if (!jc.contains("strain_indices") || !jc.contains("stress_indices"))
            throw std::runtime_error("mixed BC: strain_indices or stress_indices missing");
        bc.idx_E = jc["strain_indices"].as_int_vector();
        bc.idx_F = jc["stress_indices"].as_int_vector();
        std::vector<char> present(n_str, 0);
        for (int k : bc.idx_E) {
            if (k < 0 || k >= n_str) throw std::runtime_error("strain index out of range");
            present[k] = 1;
        }
        for (int k : bc.idx_F) {
            if (k < 0 || k >= n_str) throw std::runtime_error("stress index out of range");
            if (present[k]) throw std::runtime_error("index appears in both strain_indices and stress_indices");
            present[k] = 1;
        }
        for (int k = 0; k < n_str; ++k)
            if (!present[k]) throw std::runtime_error("each component must be either strain- or stress-controlled");
        std::vector<std::vector<double>> strain_raw, stress_raw;
        size_t n_steps = 0;
        if (!bc.idx_E.empty()) {
            if (!jc.contains("strain")) throw std::runtime_error("strain array missing");
            strain_raw = jc["strain"].as_matrix();
            n_steps = strain_raw.size();
        }
        if (!bc.idx_F.empty()) {
            if (!jc.contains("stress")) throw std::runtime_error("stress array missing");
            stress_raw = jc["stress"].as_matrix();
            n_steps = std::max(n_steps, stress_raw.size());
        }
        if (n_steps == 0) throw std::runtime_error("mixed BC: at least one of strain/stress must have timesteps");
        if (strain_raw.empty()) strain_raw.resize(n_steps);
        if (stress_raw.empty()) stress_raw.resize(n_steps);
        bc.F_E_path.assign(n_steps, std::vector<double>(bc.idx_E.size(), 0.0));
        bc.P_F_path.assign(n_steps, std::vector<double>(bc.idx_F.size(), 0.0));
        for (size_t t = 0; t < n_steps; ++t) {
            if (t >= strain_raw.size() || strain_raw[t].size() != bc.idx_E.size()) throw std::runtime_error("strain row length mismatch");
            for (size_t c = 0; c < bc.idx_E.size(); ++c) bc.F_E_path[t][c] = strain_raw[t][c];
            if (t >= stress_raw.size() || stress_raw[t].size() != bc.idx_F.size()) throw std::runtime_error("stress row length mismatch");
            for (size_t c = 0; c < bc.idx_F.size(); ++c) bc.P_F_path[t][c] = stress_raw[t][c];
        }
        bc.n_rows = n_steps;
        return bc;
    }
};

struct LoadCase {  // mixedBCs.h:139-144
    bool mixed = false;
    std::vector<std::vector<double>> g0_path;
    MixedBC mbc;
    size_t n_steps = 0;
};

}  // namespace fans
