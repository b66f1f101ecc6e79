// material_manager.hpp — mirror of include/MaterialManager.h: phase -> (model, local_mat_id, is_linear) table,
// reference stiffness, gradient broadcast.  The table is flattened into fans_phase_desc[] for the GPU library.
#pragma once
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "matmodel.hpp"
#include "reader.hpp"

namespace fans {

struct MaterialInfo {  // MaterialManager.h:13-19
    Matmodel *model = nullptr;
    LinearModelBase *linear_model = nullptr;
    int local_mat_id = -1;
    bool is_linear = false;
};

class MaterialManager {
  public:
    std::vector<std::unique_ptr<Matmodel>> models;
    std::vector<MaterialInfo> phase_to_info;
    int n_phases = 0;
    Mat kapparef_mat;
    bool all_linear = true;
    std::vector<double> g0;  // current macroscopic gradient (MaterialManager::set_gradient :216-221)
    int howmany, n_str;

    explicit MaterialManager(const Reader &reader) : howmany(reader.howmany()), n_str(reader.n_str())
    {
        // old single-matmodel input format (MaterialManager.h:42-62)
        std::vector<Json> groups;
        const Json &in = reader.inputJson;
        if (in.contains("matmodel") && in.contains("material_properties")) {
            Json g;
            g.type = Json::Object;
            g.obj.emplace_back("matmodel", in["matmodel"]);
            g.obj.emplace_back("material_properties", in["material_properties"]);
            int n_mats = 0;
            for (const auto &kv : in["material_properties"].obj)
                if (kv.second.is_array()) {
                    n_mats = (int)kv.second.size();
                    break;
                }
            Json ph;
            ph.type = Json::Array;
            for (int i = 0; i < n_mats; ++i) {
                Json v;
                v.type = Json::Number;
                v.num = i;
                ph.arr.push_back(v);
            }
            g.obj.emplace_back("phases", ph);
            groups.push_back(g);
        } else {
            const Json &mats = in.at("materials");
            if (!mats.is_array() || mats.empty()) throw std::runtime_error("MaterialManager: 'materials' must be non-empty array");
            groups = mats.arr;
        }
        int max_phase = -1;
        for (const Json &mg : groups) {
            if (!mg.contains("phases") || !mg.contains("matmodel") || !mg.contains("material_properties"))
                throw std::runtime_error("MaterialManager: material group missing required fields");
            for (int p : mg["phases"].as_int_vector()) max_phase = std::max(max_phase, p);
        }
        n_phases = max_phase + 1;
        if (n_phases == 0) throw std::runtime_error("MaterialManager: No phases defined");
        phase_to_info.assign(n_phases, MaterialInfo());
        for (const Json &mg : groups) {
            models.push_back(createMatmodel(howmany, n_str, mg["matmodel"].as_string(), mg["material_properties"], reader.ms_filename, reader.ms_datasetname));
            Matmodel *model = models.back().get();
            auto *lin = dynamic_cast<LinearModelBase *>(model);
            const bool is_linear = lin != nullptr;
            if (!is_linear) all_linear = false;
            const std::vector<int> phases = mg["phases"].as_int_vector();
            for (size_t i = 0; i < phases.size(); ++i) {
                const int p = phases[i];
                if (p < 0 || p >= n_phases || phase_to_info[p].model)
                    throw std::runtime_error("MaterialManager: Invalid or duplicate phase " + std::to_string(p));
                phase_to_info[p] = {model, lin, (int)i, is_linear};
            }
        }
        for (int p = 0; p < n_phases; ++p)
            if (!phase_to_info[p].model) throw std::runtime_error("MaterialManager: Phase " + std::to_string(p) + " not assigned");
        compute_reference_stiffness(reader);
        g0.assign(n_str, 0.0);
    }

    const MaterialInfo &get_info(int phase_id) const { return phase_to_info[phase_id]; }
    void set_gradient(const std::vector<double> &g) { g0 = g; }

    // flatten for fans_set_materials
    std::vector<fans_phase_desc> phase_descs() const
    {
        std::vector<fans_phase_desc> d(n_phases);
        for (int p = 0; p < n_phases; ++p) {
            std::fill((char *)&d[p], (char *)&d[p] + sizeof(fans_phase_desc), 0);
            const MaterialInfo &mi = phase_to_info[p];
            d[p].local_mat = mi.local_mat_id;
            d[p].group_n_mat = mi.model->n_mat;
            mi.model->fill_desc(mi.local_mat_id, d[p]);
        }
        return d;
    }

    bool has_j2() const
    {
        for (const auto &m : models)
            if (m->is_j2()) return true;
        return false;
    }

  private:
    void compute_reference_stiffness(const Reader &reader)  // MaterialManager.h:177-205
    {
        if (reader.inputJson.contains("reference_material")) {
            const auto rm = reader.inputJson["reference_material"].as_matrix();
            if ((int)rm.size() != n_str) throw std::runtime_error("reference_material must be " + std::to_string(n_str) + "x" + std::to_string(n_str));
            kapparef_mat = Mat(n_str, n_str);
            for (int i = 0; i < n_str; ++i) {
                if ((int)rm[i].size() != n_str) throw std::runtime_error("reference_material must be square");
                for (int j = 0; j < n_str; ++j) kapparef_mat(i, j) = rm[i][j];
            }
            // must be symmetric positive definite (the reference checks with an LLT factorisation)
            Mat Lc = kapparef_mat;
            for (int j = 0; j < n_str; ++j) {
                double d = Lc(j, j);
                for (int k = 0; k < j; ++k) d -= Lc(j, k) * Lc(j, k);
                if (!(d > 0.0)) throw std::runtime_error("reference_material must be symmetric positive definite");
                Lc(j, j) = std::sqrt(d);
                for (int i = j + 1; i < n_str; ++i) {
                    double s = Lc(i, j);
                    for (int k = 0; k < j; ++k) s -= Lc(i, k) * Lc(j, k);
                    Lc(i, j) = s / Lc(j, j);
                }
            }
            return;
        }
        kapparef_mat = Mat(n_str, n_str);
        for (const auto &m : models) {
            const Mat k = m->get_reference_stiffness();
            for (size_t q = 0; q < k.a.size(); ++q) kapparef_mat.a[q] += k.a[q];
        }
        for (auto &v : kapparef_mat.a) v /= (double)models.size();
    }
};

}  // namespace fans
