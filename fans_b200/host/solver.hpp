// solver.hpp — host-side Solver / SolverCG / SolverFP with the reference's public surface; all field work is
// delegated to libfans_gpu through the C ABI (include/fans_gpu.h).
//   Solver<howmany,n_str>    include/solver.h:11-103          SolverCG  include/solverCG.h:6-34
//   MixedBCController        include/mixedBCs.h:150-226       SolverFP  include/solverFP.h
//   createSolver / createMaterialManager   include/setup.h:75-90
#pragma once
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/fans_gpu.h"
#include "material_manager.hpp"
#include "reader.hpp"

namespace fans {

// Where postprocess() puts its datasets (the reference writes HDF5: reader.h:88-351).
struct ResultsSink {
    virtual ~ResultsSink() {}
    // name, load/time index, element type ("f64","f32","u16","i32"), dims = [X,Y,Z,extra...] or [n] for small data
    virtual void write(const std::string &name, int load_idx, int time_idx, const std::string &dtype, const std::vector<size_t> &dims,
                       const void *data, bool is_field) = 0;
};

class Solver {
  public:
    Reader &reader;
    MaterialManager *matmanager;
    const int world_rank, world_size;
    const ptrdiff_t n_x, n_y, n_z, local_n0, local_0_start, local_n1, local_1_start;
    const int n_it;
    double TOL;
    const int howmany, n_str;
    std::vector<double> err_all;
    size_t iter = 0;
    fans_ctx *ctx = nullptr;
    std::vector<double> homogenized_stress;
    Mat homogenized_tangent;
    std::string error_type, measure;  // reader.errorParameters (get_homogenized_tangent switches the type permanently)
    fans_solve_result last_result{};
    bool verbose = true;

    Solver(Reader &rd, MaterialManager *mm)
        : reader(rd), matmanager(mm), world_rank(rd.world_rank), world_size(rd.world_size), n_x(rd.dims[0]), n_y(rd.dims[1]), n_z(rd.dims[2]),
          local_n0(rd.local_n0), local_0_start(rd.local_0_start), local_n1(rd.local_n1), local_1_start(rd.local_1_start), n_it(rd.n_it),
          TOL(rd.TOL), howmany(rd.howmany()), n_str(rd.n_str()), error_type(rd.error_type), measure(rd.measure)
    {
        fans_config cfg{};
        for (int d = 0; d < 3; ++d) cfg.dims[d] = rd.dims[d], cfg.L[d] = rd.L[d];
        cfg.howmany = howmany;
        cfg.n_str = n_str;
        cfg.fe_type = rd.FE_type == "HEX8" ? FANS_FE_HEX8 : (rd.FE_type == "HEX8R" ? FANS_FE_HEX8R : FANS_FE_BBAR);
        cfg.world_size = world_size;
        cfg.world_rank = world_rank;
        cfg.local_n0 = (int)local_n0, cfg.local_0_start = (int)local_0_start;
        cfg.local_n1 = (int)local_n1, cfg.local_1_start = (int)local_1_start;
        cfg.device = rd.device;
        cfg.nccl_comm = rd.comm;   // one rank per GPU, rank order = slab order (main.cpp: slab_world)
        if (fans_create(&ctx, &cfg) != FANS_OK) throw std::runtime_error(std::string("fans_create: ") + fans_last_error(nullptr));
        const auto descs = matmanager->phase_descs();
        check(fans_set_materials(ctx, (int)descs.size(), descs.data()));
        check(fans_set_microstructure(ctx, rd.ms.data()));
        check(fans_set_reference_stiffness(ctx, matmanager->kapparef_mat.a.data()));  // computeFundamentalSolution
        err_all.assign(n_it + 1, 0.0);
    }
    virtual ~Solver()
    {
        if (ctx) fans_destroy(ctx);
    }
    Solver(const Solver &) = delete;

    virtual int method_id() const = 0;
    virtual void internalSolve() { run_solve(); }

    // Solver::solve, solver.h:282-300
    void solve()
    {
        std::fill(err_all.begin(), err_all.end(), 0.0);
        internalSolve();
    }

    void extrapolateDisplacement() { check(fans_extrapolate_displacement(ctx)); }  // solver.h:302-311

    std::vector<double> get_homogenized_stress()  // solver.h:707-737
    {
        push_gradient();
        homogenized_stress.assign(n_str, 0.0);
        check(fans_homogenized_stress(ctx, homogenized_stress.data()));
        return homogenized_stress;
    }

    Mat get_homogenized_tangent(double pert_param)  // solver.h:739-778
    {
        homogenized_tangent = Mat(n_str, n_str);
        const std::vector<double> unperturbed = get_homogenized_stress();
        const std::vector<double> g0 = matmanager->g0;
        const bool islinear = matmanager->all_linear;
        error_type = "relative";  // permanent, like the reference (quirk 7)
        TOL = std::max(1e-6, TOL);
        if (matmanager->has_j2()) throw std::runtime_error("Homogenized tangent computation not implemented for J2Plasticity models.");
        if (islinear && batch_tangent && tangent_batched()) return homogenized_tangent;
        for (int i = 0; i < n_str; ++i) {
            std::vector<double> pert(n_str, 0.0);
            if (islinear) {
                pert[i] = 1.0;
            } else {
                pert = g0;
                pert[i] += pert_param;
            }
            matmanager->set_gradient(pert);
            disableMixedBC();
            solve();
            const std::vector<double> p = get_homogenized_stress();
            for (int r = 0; r < n_str; ++r) homogenized_tangent(r, i) = islinear ? p[r] : (p[r] - unperturbed[r]) / pert_param;
        }
        Mat sym(n_str, n_str);
        for (int r = 0; r < n_str; ++r)
            for (int c = 0; c < n_str; ++c) sym(r, c) = 0.5 * (homogenized_tangent(r, c) + homogenized_tangent(c, r));
        homogenized_tangent = sym;
        return homogenized_tangent;
    }

    // The n_str unit load cases of the linear branch above as lanes of ONE CG loop (fans_solve_batch) instead of n_str solves one
    // after the other.  Leaves the solver in the state the reference's loop leaves it in: gradient = last unit strain, u = its
    // solution, mixed BCs off.  false: the library cannot batch this problem (slabs, FP method, odd grids) — the loop above runs.
    bool batch_tangent = true;   // FANS_TANGENT_BATCH=0 in the environment switches the batched path off (A/B runs)
    std::vector<fans_solve_result> batch_results;
    bool tangent_batched()
    {
        if (const char *env = getenv("FANS_TANGENT_BATCH"))
            if (env[0] == '0') return false;
        fans_solve_params p = solve_params();
        if (p.method != FANS_METHOD_CG) return false;
        disableMixedBC();
        std::vector<double> macro((size_t)n_str * n_str, 0.0), sig((size_t)n_str * n_str, 0.0);
        for (int i = 0; i < n_str; ++i) macro[(size_t)i * n_str + i] = 1.0;
        batch_results.assign(n_str, fans_solve_result{});
        p.verbose = 0;
        const int rc = fans_solve_batch(ctx, n_str, macro.data(), &p, batch_results.data(), sig.data(), nullptr);
        if (rc == FANS_ERR_STATE) return false;
        if (rc == FANS_ERR_CUDA && std::string(fans_last_error(ctx)).find("out of device memory") != std::string::npos) return false;
        check(rc);
        for (int i = 0; i < n_str; ++i)
            for (int r = 0; r < n_str; ++r) homogenized_tangent(r, i) = sig[(size_t)i * n_str + r];
        Mat sym(n_str, n_str);
        for (int r = 0; r < n_str; ++r)
            for (int c = 0; c < n_str; ++c) sym(r, c) = 0.5 * (homogenized_tangent(r, c) + homogenized_tangent(c, r));
        homogenized_tangent = sym;
        // state after the reference's loop (solver.h:762-775)
        std::vector<double> last(n_str, 0.0);
        last[n_str - 1] = 1.0;
        matmanager->set_gradient(last);
        push_gradient();
        check(fans_batch_load_displacement(ctx, n_str - 1, FANS_FIELD_U));
        last_result = batch_results[n_str - 1];
        iter = (size_t)last_result.iters;
        // the lane buffers are 7 fields per load case: give them back unless they are small (a micro problem that is solved again and again)
        if ((double)n_str * 7.0 * howmany * (double)local_n0 * n_y * n_z * sizeof(double) > 1e9) check(fans_batch_release(ctx));
        if (verbose && world_rank == 0) {
            int total = 0;
            for (const auto &r : batch_results) total += r.iters;
            printf("# Homogenized tangent: %d load cases in one batched CG loop, %d iterations in total, %2.6f sec\n", n_str, total,
                   1e-3 * last_result.elapsed_ms);
        }
        return true;
    }

    // ---- MixedBCController (mixedBCs.h:150-226) ----
    bool mixed_active = false;
    void enableMixedBC(const MixedBC &mbc_in, size_t t)  // activate(), mixedBCs.h:180-226
    {
        mixed_active = true;
        mbc_local = mbc_in;
        step_idx = t;
        mbc_local.finalize(matmanager->kapparef_mat);
        if (t == 0) {
            g0_vec.assign(n_str, 0.0);
            if (n_str == 9) g0_vec[0] = g0_vec[4] = g0_vec[8] = 1.0;
            g0_vec_prev = g0_vec;
        } else {
            std::vector<double> delta(n_str);
            for (int i = 0; i < n_str; ++i) delta[i] = g0_vec[i] - g0_vec_prev[i];
            g0_vec_prev = g0_vec;
            for (int k : mbc_local.idx_F) g0_vec[k] += delta[k];
        }
        for (size_t i = 0; i < mbc_local.idx_E.size(); ++i) g0_vec[mbc_local.idx_E[i]] = mbc_local.F_E_path[t][i];
        matmanager->set_gradient(g0_vec);
        push_mixed();
        updateMixedBC();
    }
    void disableMixedBC()
    {
        mixed_active = false;
        check(fans_set_mixed_bc(ctx, nullptr));
    }
    bool isMixedBCActive() const { return mixed_active; }
    void updateMixedBC()  // update(), mixedBCs.h:160-178 (runs in libfans_gpu; g0 is mirrored back)
    {
        if (!mixed_active) return;
        push_gradient();
        check(fans_update_mixed_bc(ctx));
        pull_gradient();
    }

    // Solver::postprocess, solver.h:454-705 — data sources come from the device, layout [X][Y][Z][extra]
    void postprocess(ResultsSink &sink, int load_idx, int time_idx)
    {
        const auto &results = reader.resultsToWrite;
        auto want = [&](const std::string &n) { return std::find(results.begin(), results.end(), n) != results.end(); };
        const size_t N = (size_t)local_n0 * n_y * n_z;
        push_gradient();
        // what was asked for (solver.h:459-481): ONE getStrainStress sweep, and only when something needs it — the sweep calls the
        // material law, and J2Plasticity's psi / psi_bar accumulate on every call (J2Plasticity.h:103-104)
        const int n_mat = reader.n_mat;
        const bool need_stress = want("stress"), need_strain = want("strain"), need_stress_gp = want("stress_gp"), need_strain_gp = want("strain_gp");
        const bool need_global_avg = want("stress_average") || want("strain_average");
        auto want_phase = [&](const char *base, int m) {   // the reference filters on the full name; the bare key asks for every phase
            return want(base) || want(std::string(base) + "_phase" + std::to_string(m));
        };
        bool need_phase_avg = false;
        for (int m = 0; m < n_mat; ++m) need_phase_avg = need_phase_avg || want_phase("phase_stress_average", m) || want_phase("phase_strain_average", m);
        const bool need_compute = need_stress || need_stress_gp || need_strain || need_strain_gp || need_global_avg || need_phase_avg;
        const size_t n_gp = reader.FE_type == "HEX8R" ? 1 : 8;  // matmodel.h:104-155
        std::vector<double> strain, stress, strain_gp, stress_gp;
        if (need_compute) {
            strain.resize(N * n_str), stress.resize(N * n_str);
            if (need_strain_gp) strain_gp.resize(N * n_gp * n_str);
            if (need_stress_gp) stress_gp.resize(N * n_gp * n_str);
            check(fans_strain_stress_gp(ctx, strain.data(), stress.data(), need_strain_gp ? strain_gp.data() : nullptr,
                                        need_stress_gp ? stress_gp.data() : nullptr));
        }
        warn_unknown_results(results, n_mat);
        // averages (solver.h:540-575)
        std::vector<double> sa(n_str, 0.0), ea(n_str, 0.0);
        std::vector<std::vector<double>> psa(n_mat, std::vector<double>(n_str, 0.0)), pea(n_mat, std::vector<double>(n_str, 0.0));
        std::vector<long> cnt(n_mat, 0);
        for (size_t e = 0; e < (need_compute ? N : 0); ++e) {
            const int ph = reader.ms[e];
            for (int c = 0; c < n_str; ++c) {
                sa[c] += stress[e * n_str + c];
                ea[c] += strain[e * n_str + c];
                if (ph < n_mat) psa[ph][c] += stress[e * n_str + c], pea[ph][c] += strain[e * n_str + c];
            }
            if (ph < n_mat) cnt[ph]++;
        }
        if (world_size > 1 && need_compute) {   // the MPI_Allreduce calls of solver.h:556-571, packed into one
            std::vector<double> pack;
            pack.insert(pack.end(), sa.begin(), sa.end());
            pack.insert(pack.end(), ea.begin(), ea.end());
            for (int m = 0; m < n_mat; ++m) {
                pack.insert(pack.end(), psa[m].begin(), psa[m].end());
                pack.insert(pack.end(), pea[m].begin(), pea[m].end());
                pack.push_back((double)cnt[m]);
            }
            check(fans_allreduce_sum(ctx, pack.data(), (int)pack.size()));
            size_t k = 0;
            for (int c = 0; c < n_str; ++c) sa[c] = pack[k++];
            for (int c = 0; c < n_str; ++c) ea[c] = pack[k++];
            for (int m = 0; m < n_mat; ++m) {
                for (int c = 0; c < n_str; ++c) psa[m][c] = pack[k++];
                for (int c = 0; c < n_str; ++c) pea[m][c] = pack[k++];
                cnt[m] = (long)(pack[k++] + 0.5);
            }
        }
        const double Ntot = (double)n_x * n_y * n_z;
        for (int c = 0; c < n_str; ++c) sa[c] /= Ntot, ea[c] /= Ntot;
        for (int m = 0; m < n_mat; ++m)
            if (cnt[m] > 0)
                for (int c = 0; c < n_str; ++c) psa[m][c] /= cnt[m], pea[m][c] /= cnt[m];
        stress_average = sa;
        strain_average = ea;
        if (verbose && world_rank == 0) {
            printf("# Effective Stress .. (");
            for (int i = 0; i < n_str; ++i) printf("%+.12f ", sa[i]);
            printf(") \n# Effective Strain .. (");
            for (int i = 0; i < n_str; ++i) printf("%+.12f ", ea[i]);
            printf(") \n\n");
        }
        // u_total = g0.X + u~ (solver.h:587-650)
        std::vector<double> u(N * howmany), r(N * howmany);
        check(fans_field_download(ctx, FANS_FIELD_U, u.data()));
        check(fans_field_download(ctx, FANS_FIELD_R, r.data()));
        std::vector<double> ut(u);
        const double dx = reader.l_e[0], dy = reader.l_e[1], dz = reader.l_e[2];
        const double Lx2 = reader.L[0] / 2.0, Ly2 = reader.L[1] / 2.0, Lz2 = reader.L[2] / 2.0;
        const double rs2 = 0.7071067811865475;
        size_t n = 0;
        for (ptrdiff_t ix = 0; ix < local_n0; ++ix) {
            const double X = (local_0_start + ix) * dx - Lx2;
            for (ptrdiff_t iy = 0; iy < n_y; ++iy) {
                const double Y = iy * dy - Ly2;
                for (ptrdiff_t iz = 0; iz < n_z; ++iz) {
                    const double Z = iz * dz - Lz2;
                    if (howmany == 3 && n_str == 6) {
                        const double g11 = ea[0], g22 = ea[1], g33 = ea[2], g12 = ea[3] * rs2, g13 = ea[4] * rs2, g23 = ea[5] * rs2;
                        ut[n] += g11 * X + g12 * Y + g13 * Z;
                        ut[n + 1] += g12 * X + g22 * Y + g23 * Z;
                        ut[n + 2] += g13 * X + g23 * Y + g33 * Z;
                    } else if (howmany == 3) {
                        ut[n] += (ea[0] - 1.0) * X + ea[1] * Y + ea[2] * Z;
                        ut[n + 1] += ea[3] * X + (ea[4] - 1.0) * Y + ea[5] * Z;
                        ut[n + 2] += ea[6] * X + ea[7] * Y + (ea[8] - 1.0) * Z;
                    } else {
                        ut[n] += ea[0] * X + ea[1] * Y + ea[2] * Z;
                    }
                    n += howmany;
                }
            }
        }
        const std::vector<size_t> vdim{(size_t)n_str};
        const std::vector<size_t> g{(size_t)local_n0, (size_t)n_y, (size_t)n_z};
        auto gd = [&](size_t extra) {
            std::vector<size_t> d = g;
            if (extra > 1) d.push_back(extra);
            return d;
        };
        if (want("stress_average")) sink.write("stress_average", load_idx, time_idx, "f64", vdim, sa.data(), false);
        if (want("strain_average")) sink.write("strain_average", load_idx, time_idx, "f64", vdim, ea.data(), false);
        for (int m = 0; m < n_mat; ++m) {
            if (want_phase("phase_stress_average", m)) sink.write("phase_stress_average_phase" + std::to_string(m), load_idx, time_idx, "f64", vdim, psa[m].data(), false);
            if (want_phase("phase_strain_average", m)) sink.write("phase_strain_average_phase" + std::to_string(m), load_idx, time_idx, "f64", vdim, pea[m].data(), false);
        }
        if (want("absolute_error")) sink.write("absolute_error", load_idx, time_idx, "f64", {iter + 1}, err_all.data(), false);
        if (want("microstructure")) sink.write("microstructure", load_idx, time_idx, "u16", gd(1), reader.ms.data(), true);
        if (want("displacement_fluctuation")) sink.write("displacement_fluctuation", load_idx, time_idx, "f64", gd(howmany), u.data(), true);
        if (want("displacement")) sink.write("displacement", load_idx, time_idx, "f64", gd(howmany), ut.data(), true);
        if (want("residual")) sink.write("residual", load_idx, time_idx, "f64", gd(howmany), r.data(), true);
        if (need_strain) sink.write("strain", load_idx, time_idx, "f64", gd(n_str), strain.data(), true);
        if (need_stress) sink.write("stress", load_idx, time_idx, "f64", gd(n_str), stress.data(), true);
        {   // every Gauss point, extra dims {n_gp, n_str} (solver.h:677-680)
            std::vector<size_t> d = g;
            d.push_back(n_gp);
            d.push_back((size_t)n_str);
            if (need_strain_gp) sink.write("strain_gp", load_idx, time_idx, "f64", d, strain_gp.data(), true);
            if (need_stress_gp) sink.write("stress_gp", load_idx, time_idx, "f64", d, stress_gp.data(), true);
        }
        // model postprocess (PseudoPlastic.h:55-63, J2Plasticity.h:245-322, J2PlasticityNew.h)
        auto try_field = [&](const char *nm, const char *dt, size_t extra, size_t esz) {
            if (!want(nm)) return;
            std::vector<char> buf(N * extra * esz);
            if (fans_get_field(ctx, nm, buf.data(), buf.size()) == FANS_OK) sink.write(nm, load_idx, time_idx, dt, gd(extra), buf.data(), true);
        };
        if (want("mpi_rank")) {  // solver.h:666-667
            std::vector<int> ranks(N, world_rank);
            sink.write("mpi_rank", load_idx, time_idx, "i32", gd(1), ranks.data(), true);
        }
        // all Gauss points of the committed history, extra dims {n_gp, n_str} / {n_gp} (J2Plasticity.h:298-307)
        auto try_field_gp = [&](const char *nm, size_t ncomp) {
            if (!want(nm)) return;
            const size_t ngp = reader.FE_type == "HEX8R" ? 1 : 8;  // matmodel.h:104-155: one Gauss point for HEX8R, 2x2x2 otherwise
            std::vector<double> buf(N * ngp * ncomp);
            if (fans_get_field(ctx, nm, buf.data(), buf.size() * sizeof(double)) != FANS_OK) return;
            std::vector<size_t> d = g;
            d.push_back(ngp);
            if (ncomp > 1) d.push_back(ncomp);
            sink.write(nm, load_idx, time_idx, "f64", d, buf.data(), true);
        };
        if (want("GBnormals")) {  // GBDiffusion::postprocess (GBDiffusion.h:163-175): the boundary normal of every voxel's tag
            for (const auto &mdl : matmanager->models)
                if (auto *gb = dynamic_cast<GBDiffusion *>(mdl.get())) {
                    std::vector<double> nf(N * 3, 0.0);
                    for (size_t e = 0; e < N; ++e) {
                        const size_t tag = reader.ms[e];
                        if (tag * 3 + 2 < gb->GBnormals.size())
                            for (int dcomp = 0; dcomp < 3; ++dcomp) nf[e * 3 + dcomp] = gb->GBnormals[tag * 3 + dcomp];
                    }
                    sink.write("GBnormals", load_idx, time_idx, "f64", gd(3), nf.data(), true);
                }
        }
        try_field_gp("plastic_strain_gp", 6);
        try_field_gp("isotropic_hardening_variable_gp", 1);
        try_field_gp("kinematic_hardening_variable_gp", 6);
        try_field("plastic_flag", "f32", 1, 4);
        try_field("plastic_strain", "f64", 6, 8);
        try_field("isotropic_hardening_variable", "f64", 1, 8);
        try_field("kinematic_hardening_variable", "f64", 6, 8);
        if (want("homogenized_tangent")) {
            homogenized_tangent = get_homogenized_tangent(1e-6);
            sink.write("homogenized_tangent", load_idx, time_idx, "f64", {(size_t)n_str, (size_t)n_str}, homogenized_tangent.a.data(), false);
        }
    }

    std::vector<double> stress_average, strain_average;

    // result names this front end does not produce are reported once instead of being dropped silently
    void warn_unknown_results(const std::vector<std::string> &results, int n_mat)
    {
        if (warned_results || world_rank != 0) return;
        warned_results = true;
        static const char *known[] = {"stress", "strain", "stress_gp", "strain_gp", "stress_average", "strain_average", "phase_stress_average",
                                      "phase_strain_average", "absolute_error", "microstructure", "displacement_fluctuation", "displacement",
                                      "residual", "mpi_rank", "plastic_flag", "plastic_strain", "isotropic_hardening_variable",
                                      "kinematic_hardening_variable", "plastic_strain_gp", "isotropic_hardening_variable_gp",
                                      "kinematic_hardening_variable_gp", "homogenized_tangent", "GBnormals"};
        for (const std::string &r : results) {
            bool ok = false;
            for (const char *k : known) ok = ok || r == k;
            for (int m = 0; m < n_mat && !ok; ++m)
                ok = r == "phase_stress_average_phase" + std::to_string(m) || r == "phase_strain_average_phase" + std::to_string(m);
            if (!ok) fprintf(stderr, "# WARNING: result '%s' is not produced by the GPU front end and is skipped\n", r.c_str());
        }
    }
    bool warned_results = false;

  protected:
    MixedBC mbc_local;
    size_t step_idx = 0;
    std::vector<double> g0_vec, g0_vec_prev;

    void check(int rc) const
    {
        if (rc != FANS_OK) throw std::runtime_error(fans_last_error(ctx));
    }
    void push_gradient() { check(fans_set_gradient(ctx, matmanager->g0.data())); }
    void pull_gradient()
    {
        std::vector<double> g(n_str);
        check(fans_get_gradient(ctx, g.data()));
        matmanager->g0 = g;
        if (mixed_active) g0_vec = g;
    }
    void push_mixed()
    {
        fans_mixed_bc d{};
        d.n_F = (int)mbc_local.idx_F.size();
        for (int i = 0; i < d.n_F; ++i) {
            d.idx_F[i] = mbc_local.idx_F[i];
            d.P_target[i] = mbc_local.P_F_path[step_idx][i];
            for (int j = 0; j < d.n_F; ++j) d.M[i * d.n_F + j] = mbc_local.M(i, j);
        }
        check(fans_set_mixed_bc(ctx, &d));
    }
    fans_solve_params solve_params() const
    {
        fans_solve_params p{};
        p.method = method_id();
        p.n_it = n_it;
        p.tol = TOL;
        if (measure == "L1") p.measure = FANS_MEASURE_L1;
        else if (measure == "L2") p.measure = FANS_MEASURE_L2;
        else if (measure == "Linfinity") p.measure = FANS_MEASURE_LINF;
        else throw std::runtime_error("Unknown measure type: " + measure);
        if (error_type == "absolute") p.err_type = FANS_ERR_ABSOLUTE;
        else if (error_type == "relative") p.err_type = FANS_ERR_RELATIVE;
        else throw std::runtime_error("Unknown error type: " + error_type);
        p.ls_max_iter = reader.ls_max_iter;
        p.ls_tol = reader.ls_tol;
        p.verbose = verbose && world_rank == 0;
        return p;
    }
    void run_solve()
    {
        push_gradient();
        if (mixed_active) push_mixed();
        fans_solve_params p = solve_params();
        check(fans_solve(ctx, &p, &last_result, err_all.data()));
        iter = (size_t)last_result.iters;
        pull_gradient();
        if (verbose && world_rank == 0) {  // solver.h:292-298
            const double it = (double)std::max<size_t>(iter, 1);
            printf("# FFT Time per iteration ....... %2.6f sec\n", 1e-3 * last_result.fft_ms / it);
            printf("# Total Time per iteration ..... %2.6f sec\n", 1e-3 * last_result.elapsed_ms / it);
            printf("# Total Time ................... %2.6f sec\n", 1e-3 * last_result.elapsed_ms);
        }
    }
};

class SolverCG : public Solver {  // include/solverCG.h
  public:
    using Solver::Solver;
    int method_id() const override { return FANS_METHOD_CG; }
    void internalSolve() override
    {
        if (verbose && world_rank == 0) printf("\n# Start FANS - Conjugate Gradient Solver \n");
        run_solve();
        if (verbose && world_rank == 0) printf("# Complete FANS - Conjugate Gradient Solver \n");
    }
};

class SolverFP : public Solver {  // include/solverFP.h
  public:
    using Solver::Solver;
    int method_id() const override { return FANS_METHOD_FP; }
    void internalSolve() override
    {
        if (verbose && world_rank == 0) printf("\n# Start FANS - Fixed Point Iteration Solver \n");
        run_solve();
        if (verbose && world_rank == 0) printf("# Complete FANS - Fixed Point Iteration Solver \n");
    }
};

// include/setup.h:75-90
static inline Solver *createSolver(Reader &reader, MaterialManager *matmanager)
{
    if (reader.method == "fp") return new SolverFP(reader, matmanager);
    if (reader.method == "cg") return new SolverCG(reader, matmanager);
    throw std::invalid_argument(reader.method + " is not a valid method");
}
static inline MaterialManager *createMaterialManager(const Reader &reader) { return new MaterialManager(reader); }

}  // namespace fans
