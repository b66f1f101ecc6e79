// pyfans.cpp — the `PyFANS` Python module for two-scale coupling through the preCICE Micro Manager, GPU build.
// Same surface as the reference's pybind11 module (pyfans/micro.hpp:20-32, pyfans/micro.cpp:27-95):
//     sim = PyFANS.MicroSimulation(sim_id)               # reads "input.json" (mechanical, small strain: <howmany, n_str> = <3, 6>)
//     out = sim.solve({"strains1to3": [...], "strains4to6": [...]}, dt)
//     out -> {"stresses1to3", "stresses4to6", "cmat1" .. "cmat7"}   (upper triangle of the homogenized tangent, three at a time)
// The solve, the homogenized stress and the n_str tangent solves all run in libfans_gpu through the C ABI.
#include <memory>

#include "pybind11/numpy.h"
#include "pybind11/pybind11.h"
#include "pybind11/stl.h"
#include "solver.hpp"

namespace py = pybind11;

class MicroSimulation {
  public:
    explicit MicroSimulation(int sim_id, const std::string &input_file = "input.json") : sim_id_(sim_id)
    {
        reader_.ReadInputFile(input_file);
        if (reader_.howmany() != 3 || reader_.n_str() != 6)
            throw std::runtime_error("PyFANS.MicroSimulation supports mechanical small-strain inputs only (like the reference)");
        reader_.ReadMS(3);
        matmanager_.reset(fans::createMaterialManager(reader_));
        solver_.reset(fans::createSolver(reader_, matmanager_.get()));
        solver_->verbose = false;
    }

    // micro.cpp:40-85.  dt is accepted and unused, like in the reference.
    py::dict solve(py::dict macro_data, double /*dt*/)
    {
        std::vector<double> g0;
        for (const char *key : {"strains1to3", "strains4to6"}) {
            auto part = py::array_t<double, py::array::c_style | py::array::forcecast>::ensure(macro_data[key]);
            if (!part) throw std::invalid_argument(std::string("macro data entry '") + key + "' is not a float array");
            g0.insert(g0.end(), part.data(), part.data() + part.size());
        }
        if (g0.size() != 6) throw std::invalid_argument("expected 6 strain components in strains1to3 + strains4to6");
        matmanager_->set_gradient(g0);
        solver_->solve();
        const std::vector<double> s = solver_->get_homogenized_stress();
        const fans::Mat C = solver_->get_homogenized_tangent(pert_param_);
        py::dict out;
        out["stresses1to3"] = std::vector<double>{s[0], s[1], s[2]};
        out["stresses4to6"] = std::vector<double>{s[3], s[4], s[5]};
        std::vector<double> tri;  // row-major upper triangle: 21 entries -> cmat1..cmat7
        for (int i = 0; i < 6; ++i)
            for (int j = i; j < 6; ++j) tri.push_back(C(i, j));
        for (int k = 0; k < 7; ++k) out[("cmat" + std::to_string(k + 1)).c_str()] = std::vector<double>{tri[3 * k], tri[3 * k + 1], tri[3 * k + 2]};
        return out;
    }

  private:
    int sim_id_;
    fans::Reader reader_;
    std::unique_ptr<fans::MaterialManager> matmanager_;
    std::unique_ptr<fans::Solver> solver_;
    double pert_param_ = 1e-6;  // micro.hpp:31
};

PYBIND11_MODULE(PyFANS, m)
{
    m.doc() = "FANS for Micro Manager (B200 build)";
    py::class_<MicroSimulation>(m, "MicroSimulation")
        .def(py::init<int>())
        .def(py::init<int, const std::string &>())
        .def("solve", &MicroSimulation::solve);
}
