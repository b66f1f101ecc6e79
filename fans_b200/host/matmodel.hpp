// matmodel.hpp — host-side mirror of the reference's material plug-in interface.
//   Matmodel<howmany,n_str>           include/matmodel.h:10-75
//   LinearModel<howmany,n_str>        include/matmodel.h:306-310
//   createMatmodel (string registry)  include/setup.h:21-73
// The host classes only parse `material_properties`, pre-compute the same per-material constants the reference
// constructors compute, and export one POD fans_phase_desc per local material; get_sigma itself runs on the GPU
// (fans_b200/csrc/materials.cuh).  Errors are C++ exceptions with the reference's messages.
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <memory>
#include <numeric>
#include <stdexcept>
#include <string>
#include <vector>

#include "h5mini.hpp"

#include "../../include/fans_gpu.h"
#include "json.hpp"

namespace fans {

using std::string;
using std::vector;

// dense row-major matrix, tiny
struct Mat {
    int r = 0, c = 0;
    vector<double> a;
    Mat() {}
    Mat(int r_, int c_) : r(r_), c(c_), a((size_t)r_ * c_, 0.0) {}
    double &operator()(int i, int j) { return a[(size_t)i * c + j]; }
    double operator()(int i, int j) const { return a[(size_t)i * c + j]; }
    static Mat identity(int n)
    {
        Mat m(n, n);
        for (int i = 0; i < n; ++i) m(i, i) = 1.0;
        return m;
    }
};

static inline vector<double> prop_vector(const Json &props, const string &key)
{
    try {
        return props.at(key).as_vector();
    } catch (const std::exception &) {
        throw std::runtime_error("Missing material properties for the requested material model.");
    }
}

class Matmodel {
  public:
    int howmany, num_str;
    int n_mat = 0;
    virtual ~Matmodel() {}
    Matmodel(int h, int n) : howmany(h), num_str(n) {}
    virtual bool is_linear() const { return false; }
    virtual Mat get_reference_stiffness() const = 0;
    virtual void fill_desc(int mat_index, fans_phase_desc &d) const = 0;
    virtual bool is_j2() const { return false; }
};

// ---------------- linear models: export the phase tangent, the GPU builds phase_stiffness ----------------
class LinearModelBase : public Matmodel {
  public:
    using Matmodel::Matmodel;
    bool is_linear() const override { return true; }
    virtual Mat phase_kappa(int i) const = 0;
    void fill_desc(int i, fans_phase_desc &d) const override
    {
        d.model = FANS_MAT_LINEAR;
        const Mat k = phase_kappa(i);
        for (int q = 0; q < num_str * num_str; ++q) d.params[q] = k.a[q];
    }
};

class LinearThermalIsotropic : public LinearModelBase {  // LinearThermal.h:7-46
  public:
    vector<double> conductivity;
    explicit LinearThermalIsotropic(const Json &props) : LinearModelBase(1, 3)
    {
        conductivity = prop_vector(props, "conductivity");
        n_mat = (int)conductivity.size();
    }
    Mat phase_kappa(int i) const override
    {
        Mat k = Mat::identity(3);
        for (auto &v : k.a) v *= conductivity[i];
        return k;
    }
    Mat get_reference_stiffness() const override
    {
        Mat k(3, 3);
        for (int i = 0; i < n_mat; ++i)
            for (int d = 0; d < 3; ++d) k(d, d) += conductivity[i];
        for (auto &v : k.a) v /= n_mat;
        return k;
    }
};

class LinearThermalTriclinic : public LinearModelBase {  // LinearThermal.h:48-118
  public:
    vector<Mat> K_mats;
    explicit LinearThermalTriclinic(const Json &props) : LinearModelBase(1, 3)
    {
        const char *keys[6] = {"K_11", "K_12", "K_13", "K_22", "K_23", "K_33"};
        vector<vector<double>> c;
        try {
            n_mat = (int)props.at("K_11").as_vector().size();
            for (auto k : keys) {
                c.push_back(props.at(k).as_vector());
                if ((int)c.back().size() != n_mat) throw std::runtime_error("Inconsistent size for material property: " + string(k));
            }
        } catch (const std::exception &) {
            throw std::runtime_error("Missing or inconsistent material properties for the requested material model.");
        }
        for (int i = 0; i < n_mat; ++i) {
            Mat K(3, 3);
            K(0, 0) = c[0][i], K(0, 1) = c[1][i], K(0, 2) = c[2][i];
            K(1, 0) = c[1][i], K(1, 1) = c[3][i], K(1, 2) = c[4][i];
            K(2, 0) = c[2][i], K(2, 1) = c[4][i], K(2, 2) = c[5][i];
            K_mats.push_back(K);
        }
    }
    Mat phase_kappa(int i) const override { return K_mats[i]; }
    Mat get_reference_stiffness() const override
    {
        Mat k(3, 3);
        for (const auto &m : K_mats)
            for (int q = 0; q < 9; ++q) k.a[q] += m.a[q];
        for (auto &v : k.a) v /= n_mat;
        return k;
    }
};

// GBDiffusion (GBDiffusion.h:45-175): polycrystal diffusion, a LinearModel — isotropic D_bulk in the crystals (tags 0..num_crystals-1),
// transversely isotropic D_par (I - N N^T) + D_perp N N^T in the grain-boundary phases.  num_crystals, num_GB and the boundary normals
// (GBVoxelInfo, a JSON text) are attributes of the microstructure dataset (GBDiffusion.h:51-69); images without HDF5 attributes
// (.npy / raw) may carry the same three entries as "num_crystals", "num_GB", "GBVoxelInfo" inside material_properties (extension).
// On the device this is FANS_MAT_LINEAR with one tangent per tag (the stencil's coefficient-table path takes any number of them).
// Note: the reference evaluates get_sigma with the normals as stored and phase_stiffness with the normalised ones
// (GBDiffusion.h:109-111 vs :137-161); unit normals (what MSUtils writes) make the two agree, and only those are supported here.
class GBDiffusion : public LinearModelBase {
  public:
    long long num_crystals = 0, num_GB = 0;
    vector<double> GBnormals, D_bulk, D_par, D_perp;
    GBDiffusion(const Json &props, const string &ms_file, const string &ms_dataset) : LinearModelBase(1, 3)
    {
        try {
            Json info;
            H5Attrs at;
            string err;
            const bool have = !ms_file.empty() && h5mini_read_attributes(ms_file, ms_dataset, at, err) && at.ints.count("num_crystals") &&
                              at.ints.count("num_GB") && at.strs.count("GBVoxelInfo");
            if (have) {
                num_crystals = at.ints["num_crystals"], num_GB = at.ints["num_GB"];
                info = Json::parse(at.strs["GBVoxelInfo"]);
            } else if (props.contains("num_crystals") && props.contains("num_GB") && props.contains("GBVoxelInfo")) {
                num_crystals = props.at("num_crystals").as_int(), num_GB = props.at("num_GB").as_int();
                info = props.at("GBVoxelInfo");
            } else {
                throw std::runtime_error("attributes num_crystals / num_GB / GBVoxelInfo not found on the microstructure dataset" +
                                         (err.empty() ? string() : " (" + err + ")"));
            }
            n_mat = (int)(num_crystals + num_GB);
            GBnormals.assign((size_t)n_mat * 3, 0.0);
            for (const auto &kv : info.obj) {
                const int tag = kv.second.at("GB_tag").as_int();
                const vector<double> nrm = kv.second.at("GB_normal").as_vector();
                if (tag < 0 || tag >= n_mat || nrm.size() != 3) throw std::runtime_error("bad GB_tag / GB_normal entry");
                for (int d = 0; d < 3; ++d) GBnormals[(size_t)tag * 3 + d] = nrm[d];
            }
            D_bulk.assign(n_mat, 0.0), D_par.assign(n_mat, 0.0), D_perp.assign(n_mat, 0.0);
            if (props.at("GB_unformity").as_bool()) {
                std::fill_n(D_bulk.begin(), num_crystals, props.at("D_bulk").as_double());
                std::fill_n(D_par.begin() + num_crystals, num_GB, props.at("D_par").as_double());
                std::fill_n(D_perp.begin() + num_crystals, num_GB, props.at("D_perp").as_double());
            } else {
                const vector<double> b = props.at("D_bulk").as_vector(), pa = props.at("D_par").as_vector(), pe = props.at("D_perp").as_vector();
                for (int i = 0; i < n_mat; ++i) D_bulk[i] = b.at(i), D_par[i] = pa.at(i), D_perp[i] = pe.at(i);
            }
        } catch (const std::exception &e) {
            throw std::runtime_error("Error in GBDiffusion initialization: " + string(e.what()));
        }
    }
    Mat phase_kappa(int i) const override
    {
        Mat k = Mat::identity(3);
        if (i < num_crystals) {
            for (auto &v : k.a) v *= D_bulk[i];
            return k;
        }
        double n[3] = {GBnormals[3 * i], GBnormals[3 * i + 1], GBnormals[3 * i + 2]};
        const double len = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
        if (len > 0)
            for (double &c : n) c /= len;
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) k(r, c) = D_par[i] * ((r == c ? 1.0 : 0.0) - n[r] * n[c]) + D_perp[i] * n[r] * n[c];
        return k;
    }
    Mat get_reference_stiffness() const override  // kappa_average, GBDiffusion.h:120-122
    {
        Mat k(3, 3);
        for (int i = 0; i < n_mat; ++i) {
            const Mat m = phase_kappa(i);
            for (int q = 0; q < 9; ++q) k.a[q] += m.a[q];
        }
        for (auto &v : k.a) v /= n_mat;
        return k;
    }
};

static inline Mat iso_tangent(double lambda, double mu)
{
    Mat k(6, 6);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) k(i, j) = lambda;
    for (int i = 0; i < 6; ++i) k(i, i) += 2 * mu;
    return k;
}

class LinearElasticIsotropic : public LinearModelBase {  // LinearElastic.h:7-75
  public:
    vector<double> bulk_modulus, mu, lambda;
    explicit LinearElasticIsotropic(const Json &props) : LinearModelBase(3, 6)
    {
        bulk_modulus = prop_vector(props, "bulk_modulus");
        mu = prop_vector(props, "shear_modulus");
        n_mat = (int)bulk_modulus.size();
        lambda.resize(n_mat);
        mu.resize(n_mat);
        for (int i = 0; i < n_mat; ++i) lambda[i] = bulk_modulus[i] - (2.0 / 3.0) * mu[i];
    }
    Mat phase_kappa(int i) const override { return iso_tangent(lambda[i], mu[i]); }
    Mat get_reference_stiffness() const override  // LinearElastic.h:55-69: (max+min)/2
    {
        const double lambda_ref = (*std::max_element(lambda.begin(), lambda.end()) + *std::min_element(lambda.begin(), lambda.end())) / 2;
        const double mu_ref = (*std::max_element(mu.begin(), mu.end()) + *std::min_element(mu.begin(), mu.end())) / 2;
        return iso_tangent(lambda_ref, mu_ref);
    }
};

class LinearElasticTriclinic : public LinearModelBase {  // LinearElastic.h:77-159
  public:
    vector<Mat> C_mats;
    explicit LinearElasticTriclinic(const Json &props) : LinearModelBase(3, 6)
    {
        vector<vector<double>> c;
        try {
            n_mat = (int)props.at("C_11").as_vector().size();
            for (int r = 0; r < 6; ++r)
                for (int q = r; q < 6; ++q) {
                    const string k = "C_" + std::to_string(r + 1) + std::to_string(q + 1);
                    c.push_back(props.at(k).as_vector());
                    if ((int)c.back().size() != n_mat) throw std::runtime_error("Inconsistent size for material property: " + k);
                }
        } catch (const std::exception &) {
            throw std::runtime_error("Missing or inconsistent material properties for the requested material model.");
        }
        for (int i = 0; i < n_mat; ++i) {
            Mat C(6, 6);
            int k = 0;
            for (int r = 0; r < 6; ++r)
                for (int q = r; q < 6; ++q) {
                    C(r, q) = c[k][i];
                    C(q, r) = c[k][i];
                    ++k;
                }
            C_mats.push_back(C);
        }
    }
    Mat phase_kappa(int i) const override { return C_mats[i]; }
    Mat get_reference_stiffness() const override
    {
        Mat k(6, 6);
        for (const auto &m : C_mats)
            for (int q = 0; q < 36; ++q) k.a[q] += m.a[q];
        for (auto &v : k.a) v /= n_mat;
        return k;
    }
};

// arithmetic-mean isotropic reference (PseudoPlastic.h:43-53, J2Plasticity.h:113-123, J2PlasticityNew.h)
static inline Mat mean_iso_reference(const vector<double> &K, const vector<double> &G)
{
    const double Kbar = std::accumulate(K.begin(), K.end(), 0.0) / (double)K.size();
    const double Gbar = std::accumulate(G.begin(), G.end(), 0.0) / (double)G.size();
    return iso_tangent(Kbar - 2.0 * Gbar / 3.0, Gbar);
}

class PseudoPlastic : public Matmodel {  // PseudoPlastic.h:22-76
  public:
    vector<double> bulk_modulus, shear_modulus, yield_stress, eps_crit;
    explicit PseudoPlastic(const Json &props) : Matmodel(3, 6)
    {
        bulk_modulus = prop_vector(props, "bulk_modulus");
        shear_modulus = prop_vector(props, "shear_modulus");
        yield_stress = prop_vector(props, "yield_stress");
        n_mat = (int)bulk_modulus.size();
    }
    Mat get_reference_stiffness() const override { return mean_iso_reference(bulk_modulus, shear_modulus); }
};

class PseudoPlasticLinearHardening : public PseudoPlastic {  // PseudoPlastic.h:78-124
  public:
    vector<double> hardening_parameter, E_s;
    explicit PseudoPlasticLinearHardening(const Json &props) : PseudoPlastic(props)
    {
        hardening_parameter = prop_vector(props, "hardening_parameter");
        E_s.resize(n_mat);
        eps_crit.resize(n_mat);
        for (int i = 0; i < n_mat; ++i) {
            eps_crit[i] = std::sqrt(2. / 3.) * yield_stress[i] / (2. * shear_modulus[i]);
            E_s[i] = (3. * shear_modulus[i]) / (3. * shear_modulus[i] + hardening_parameter[i]);
        }
    }
    void fill_desc(int i, fans_phase_desc &d) const override
    {
        d.model = FANS_MAT_PSEUDOPLASTIC_LINEAR;
        const double p[6] = {bulk_modulus[i], shear_modulus[i], yield_stress[i], hardening_parameter[i], eps_crit[i], E_s[i]};
        std::copy(p, p + 6, d.params);
    }
};

class PseudoPlasticNonLinearHardening : public PseudoPlastic {  // PseudoPlastic.h:126-173
  public:
    vector<double> hardening_exponent, eps_0;
    explicit PseudoPlasticNonLinearHardening(const Json &props) : PseudoPlastic(props)
    {
        hardening_exponent = prop_vector(props, "hardening_exponent");
        eps_0 = prop_vector(props, "eps_0");
        eps_crit.resize(n_mat);
        for (int i = 0; i < n_mat; ++i)
            eps_crit[i] = eps_0[i] * std::pow(yield_stress[i] / (3.0 * shear_modulus[i] * eps_0[i]), 1.0 / (1.0 - hardening_exponent[i]));
    }
    void fill_desc(int i, fans_phase_desc &d) const override
    {
        d.model = FANS_MAT_PSEUDOPLASTIC_NONLIN;
        const double p[6] = {bulk_modulus[i], shear_modulus[i], yield_stress[i], hardening_exponent[i], eps_0[i], eps_crit[i]};
        std::copy(p, p + 6, d.params);
    }
};

class J2Plasticity : public Matmodel {  // J2Plasticity.h:7-163
  public:
    vector<double> bulk_modulus, shear_modulus, yield_stress, K, H, eta;
    double dt = 0.0;
    explicit J2Plasticity(const Json &props) : Matmodel(3, 6)
    {
        bulk_modulus = prop_vector(props, "bulk_modulus");
        shear_modulus = prop_vector(props, "shear_modulus");
        yield_stress = prop_vector(props, "yield_stress");
        K = prop_vector(props, "isotropic_hardening_parameter");
        H = prop_vector(props, "kinematic_hardening_parameter");
        eta = prop_vector(props, "viscosity");
        try {
            dt = props.at("time_step").as_double();
        } catch (const std::exception &) {
            throw std::runtime_error("Missing material properties for the requested material model.");
        }
        n_mat = (int)bulk_modulus.size();
    }
    bool is_j2() const override { return true; }
    Mat get_reference_stiffness() const override { return mean_iso_reference(bulk_modulus, shear_modulus); }
};

class J2ViscoPlastic_LinearIsotropicHardening : public J2Plasticity {  // J2Plasticity.h:165-178
  public:
    using J2Plasticity::J2Plasticity;
    void fill_desc(int i, fans_phase_desc &d) const override
    {
        d.model = FANS_MAT_J2_LINEAR_ISO;
        const double p[7] = {bulk_modulus[i], shear_modulus[i], yield_stress[i], K[i], H[i], eta[i], dt};
        std::copy(p, p + 7, d.params);
    }
};

class J2ViscoPlastic_NonLinearIsotropicHardening : public J2Plasticity {  // J2Plasticity.h:180-243
  public:
    vector<double> sigma_inf, delta;
    explicit J2ViscoPlastic_NonLinearIsotropicHardening(const Json &props) : J2Plasticity(props)
    {
        sigma_inf = prop_vector(props, "saturation_stress");
        delta = prop_vector(props, "saturation_exponent");
    }
    void fill_desc(int i, fans_phase_desc &d) const override
    {
        d.model = FANS_MAT_J2_NONLIN_ISO;
        const double p[9] = {bulk_modulus[i], shear_modulus[i], yield_stress[i], K[i], H[i], eta[i], dt, sigma_inf[i], delta[i]};
        std::copy(p, p + 9, d.params);
    }
};

class J2PlasticityNew_LinearIsotropicHardening : public Matmodel {  // J2PlasticityNew.h:7-150
  public:
    vector<double> bulk_modulus, shear_modulus, yield_stress, K;
    explicit J2PlasticityNew_LinearIsotropicHardening(const Json &props) : Matmodel(3, 6)
    {
        bulk_modulus = prop_vector(props, "bulk_modulus");
        shear_modulus = prop_vector(props, "shear_modulus");
        yield_stress = prop_vector(props, "yield_stress");
        K = prop_vector(props, "isotropic_hardening_parameter");
        n_mat = (int)bulk_modulus.size();
    }
    Mat get_reference_stiffness() const override { return mean_iso_reference(bulk_modulus, shear_modulus); }
    void fill_desc(int i, fans_phase_desc &d) const override
    {
        d.model = FANS_MAT_J2NEW_LINEAR_ISO;
        const double p[4] = {bulk_modulus[i], shear_modulus[i], yield_stress[i], K[i]};
        std::copy(p, p + 4, d.params);
    }
};

// ---------------- finite strain (LargeStrainMechModel.h) ----------------
// Reference medium of the finite-strain models: A = dP/dF evaluated by LargeStrainMechModel.h:105-180 at F = I, S = 0 with the
// isotropic material tangent C = lambda 1x1 + 2 mu I_sym, in closed form.  The reference sums dE_PQ/dF_kL over P <= Q only
// (LargeStrainMechModel.h:146-147), so a shear pair (k != L) picks up HALF the tensor component C_iJkL = mu:
//   A(3i+i, 3k+k) = lambda + 2 mu delta_ik ;   A(3i+J, 3i+J) = A(3i+J, 3J+i) = mu / 2   (i != J) ;   everything else 0.
// This fixes the Green operator and with it the iteration counts, so the halved shear entries are kept on purpose.
static inline Mat spatial_tangent_at_identity(double lambda, double mu)
{
    Mat A(9, 9);
    for (int i = 0; i < 3; ++i)
        for (int k = 0; k < 3; ++k) {
            A(3 * i + i, 3 * k + k) = lambda + (i == k ? 2.0 * mu : 0.0);
            if (i != k) A(3 * i + k, 3 * i + k) = A(3 * i + k, 3 * k + i) = 0.5 * mu;
        }
    return A;
}

class LargeStrainMechModel : public Matmodel {
  public:
    vector<double> bulk_modulus, shear_modulus, lambda, mu;
    explicit LargeStrainMechModel(const Json &props) : Matmodel(3, 9)
    {
        bulk_modulus = prop_vector(props, "bulk_modulus");
        shear_modulus = prop_vector(props, "shear_modulus");
        n_mat = (int)bulk_modulus.size();
        lambda.resize(n_mat);
        mu.resize(n_mat);
        for (int i = 0; i < n_mat; ++i) {
            mu[i] = shear_modulus[i];
            lambda[i] = bulk_modulus[i] - (2.0 / 3.0) * mu[i];
        }
    }
    // both hyperelastic laws of the reference have S(F=I) = 0 and the isotropic tangent at F = I
    // (SaintVenantKirchhoff.h:40-66; CompressibleNeoHookean.h:50-108 with C^-1 = I, log J = 0:
    //  lambda PP1 + mu PP2 = lambda 1x1 + 2 mu I in Mandel notation)
    Mat get_reference_stiffness() const override
    {
        Mat kapparef(9, 9);
        for (int m = 0; m < n_mat; ++m) {
            const Mat A = spatial_tangent_at_identity(lambda[m], mu[m]);
            for (int q = 0; q < 81; ++q) kapparef.a[q] += A.a[q];
        }
        for (auto &v : kapparef.a) v /= (double)n_mat;
        return kapparef;
    }
};

class SaintVenantKirchhoff : public LargeStrainMechModel {  // SaintVenantKirchhoff.h
  public:
    using LargeStrainMechModel::LargeStrainMechModel;
    void fill_desc(int i, fans_phase_desc &d) const override
    {
        d.model = FANS_MAT_SVK;
        d.params[0] = lambda[i];
        d.params[1] = mu[i];
    }
};

class CompressibleNeoHookean : public LargeStrainMechModel {  // CompressibleNeoHookean.h
  public:
    using LargeStrainMechModel::LargeStrainMechModel;
    void fill_desc(int i, fans_phase_desc &d) const override
    {
        d.model = FANS_MAT_NEOHOOKE;
        d.params[0] = lambda[i];
        d.params[1] = mu[i];
    }
};

// createMatmodel: include/setup.h:21-73
static inline std::unique_ptr<Matmodel> createMatmodel(int howmany, int n_str, const string &name, const Json &props,
                                                       const string &ms_file = "", const string &ms_dataset = "")
{
    if (howmany == 1 && n_str == 3) {
        if (name == "LinearThermalIsotropic") return std::make_unique<LinearThermalIsotropic>(props);
        if (name == "LinearThermalTriclinic") return std::make_unique<LinearThermalTriclinic>(props);
        if (name == "GBDiffusion") return std::make_unique<GBDiffusion>(props, ms_file, ms_dataset);
        throw std::invalid_argument(name + " is not a valid matmodel for thermal problem");
    }
    if (howmany == 3 && n_str == 6) {
        if (name == "LinearElasticIsotropic") return std::make_unique<LinearElasticIsotropic>(props);
        if (name == "LinearElasticTriclinic") return std::make_unique<LinearElasticTriclinic>(props);
        if (name == "PseudoPlasticLinearHardening") return std::make_unique<PseudoPlasticLinearHardening>(props);
        if (name == "PseudoPlasticNonLinearHardening") return std::make_unique<PseudoPlasticNonLinearHardening>(props);
        if (name == "J2ViscoPlastic_LinearIsotropicHardening") return std::make_unique<J2ViscoPlastic_LinearIsotropicHardening>(props);
        if (name == "J2ViscoPlastic_NonLinearIsotropicHardening") return std::make_unique<J2ViscoPlastic_NonLinearIsotropicHardening>(props);
        if (name == "J2PlasticityNew_LinearIsotropicHardening") return std::make_unique<J2PlasticityNew_LinearIsotropicHardening>(props);
        throw std::invalid_argument(name + " is not a valid small strain material model");
    }
    if (howmany == 3 && n_str == 9) {
        if (name == "SaintVenantKirchhoff") return std::make_unique<SaintVenantKirchhoff>(props);
        if (name == "CompressibleNeoHookean") return std::make_unique<CompressibleNeoHookean>(props);
        throw std::invalid_argument(name + " is not a valid large strain material model");
    }
    throw std::invalid_argument("invalid (howmany, n_str)");
}

}  // namespace fans
