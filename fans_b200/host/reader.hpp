// reader.hpp — input side of the host: mirror of the reference Reader (include/reader.h, src/reader.cpp).
//   ReadInputFile   src/reader.cpp:63-183     (same JSON keys, defaults and error messages)
//   ReadMS          src/reader.cpp:227-411    (microstructure -> memory order [x][y][z], slab sizes, volume fractions)
// HDF5 is not available in this image: microstructures come from a minimal built-in HDF5 reader (h5mini.hpp),
// from .npy files, or are handed over in memory (SetMicrostructure) by an embedding application.
#pragma once
#include <cstdio>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "h5mini.hpp"
#include "json.hpp"
#include "mixedbc.hpp"

namespace fans {


class Reader {
  public:
    // contents of input file (src/reader.cpp:63-183)
    std::string ms_filename, ms_datasetname, results_prefix, dataset_name;
    int n_it = 0;
    double TOL = 0.0;
    std::string measure, error_type;  // errorParameters["measure"], ["type"]
    Json inputJson, microstructure;
    std::string problemType, matmodel, method, strain_type = "small", FE_type = "HEX8";
    std::vector<std::string> resultsToWrite;
    std::vector<LoadCase> load_cases;
    bool extrapolate_displacement = true;  // include/reader.h:35
    int ls_max_iter = 5;                   // include/reader.h (linesearch defaults 5 / 1e-2)
    double ls_tol = 1e-2;

    // contents of microstructure file
    std::vector<int> dims;      // n_x, n_y, n_z
    std::vector<double> l_e, L;
    std::vector<uint16_t> ms;   // [x][y][z]
    int n_mat = 0;

    int world_rank = 0, world_size = 1;
    int local_n0 = 0, local_0_start = 0, local_n1 = 0, local_1_start = 0;
    void *comm = nullptr;   // slab communicator (fans_comm_create) when world_size > 1, see main.cpp
    int device = -1;        // CUDA device of this rank (-1: current)

    int howmany() const { return problemType == "thermal" ? 1 : 3; }
    int n_str() const { return problemType == "thermal" ? 3 : (strain_type == "large" ? 9 : 6); }

    void ReadInputFile(const std::string &fn)
    {
        std::ifstream f(fn);
        if (!f) throw std::runtime_error("cannot open input file '" + fn + "'");
        std::stringstream ss;
        ss << f.rdbuf();
        ReadInputString(ss.str());
    }

    void ReadInputString(const std::string &text)
    {
        const Json j = Json::parse(text);
        inputJson = j;
        microstructure = j["microstructure"];
        ms_filename = microstructure.contains("filepath") ? microstructure["filepath"].as_string() : "";
        const std::string tmp = microstructure["datasetname"].as_string();
        if (tmp.empty()) throw std::invalid_argument("datasetname must not be empty and must refer to a valid HDF5 path");
        ms_datasetname = (tmp.front() == '/' ? "" : "/") + tmp;
        L = microstructure["L"].as_vector();
        results_prefix = j.contains("results_prefix") ? j["results_prefix"].as_string() : "";
        dataset_name = ms_datasetname + "_results/" + results_prefix;

        const Json &ep = j["error_parameters"];
        TOL = ep["tolerance"].as_double();
        measure = ep["measure"].as_string();
        error_type = ep["type"].as_string();
        n_it = j["n_it"].as_int();
        extrapolate_displacement = j.value("extrapolate_displacement", extrapolate_displacement);
        if (j.contains("linesearch_parameters")) {
            ls_max_iter = j["linesearch_parameters"].value("max_iter", ls_max_iter);
            ls_tol = j["linesearch_parameters"].value("tol", ls_tol);
            if (ls_max_iter < 1 || ls_tol <= 0.0) throw std::invalid_argument("linesearch_parameters: max_iter >= 1 and tol > 0 required");
        }
        problemType = j["problem_type"].as_string();
        method = j["method"].as_string();
        if (j.contains("strain_type")) {
            strain_type = j["strain_type"].as_string();
            if (strain_type != "small" && strain_type != "large") throw std::invalid_argument("strain_type must be either 'small' or 'large'");
        } else {
            strain_type = "small";
        }
        if (j.contains("FE_type")) {
            FE_type = j["FE_type"].as_string();
            if (FE_type != "HEX8" && FE_type != "HEX8R" && FE_type != "BBAR")
                throw std::invalid_argument("FE_type must be one of: 'HEX8', 'HEX8R', or 'BBAR'");
        } else {
            FE_type = "HEX8";
        }
        if (problemType != "thermal" && problemType != "mechanical") throw std::invalid_argument(problemType + " is not a valid problem type");
        resultsToWrite = j["results"].as_string_vector();

        load_cases.clear();
        const Json &ml = j["macroscale_loading"];
        if (!ml.is_array()) throw std::runtime_error("macroscale_loading must be an array");
        const int ns = n_str();
        for (const Json &entry : ml.arr) {
            LoadCase lc;
            if (entry.is_array()) {  // legacy pure-strain
                lc.mixed = false;
                lc.g0_path = entry.as_matrix();
                lc.n_steps = lc.g0_path.size();
                if (lc.g0_path.empty() || lc.g0_path[0].size() != (size_t)ns)
                    throw std::invalid_argument("Invalid length of loading vector: expected " + std::to_string(ns) + " components but got " +
                                                std::to_string(lc.g0_path.empty() ? 0 : lc.g0_path[0].size()));
            } else {
                lc.mixed = true;
                lc.mbc = MixedBC::from_json(entry, ns);
                lc.n_steps = lc.mbc.n_rows;
            }
            load_cases.push_back(std::move(lc));
        }
    }

    // microstructure handed over in memory, already in the solver's [x][y][z] order
    void SetMicrostructure(const int d[3], const uint16_t *data)
    {
        dims.assign(d, d + 3);
        ms.assign(data, data + (size_t)d[0] * d[1] * d[2]);
        finish_ms();
    }

    // src/reader.cpp:227-411.  On disk the array is [z][y][x] unless the attribute permute_order says "xyz".
    void ReadMS(int /*hm*/)
    {
        if (!ms.empty()) return;  // provided in memory
        std::vector<int> fdims;
        std::vector<uint16_t> raw;
        std::string order = "zyx", err;
        const bool is_npy = ms_filename.size() > 4 && ms_filename.compare(ms_filename.size() - 4, 4, ".npy") == 0;
        if (!(is_npy ? npy_read(ms_filename, fdims, raw, err) : h5mini_read_dataset(ms_filename, ms_datasetname, fdims, raw, order, err)))
            throw std::runtime_error("cannot read microstructure '" + ms_filename + "':'" + ms_datasetname + "': " + err);
        if (is_npy && microstructure.contains("permute_order")) order = microstructure["permute_order"].as_string();
        if (fdims.size() != 3) throw std::runtime_error("microstructure dataset must be 3-dimensional");
        if (order == "xyz") {
            dims = fdims;
            ms = raw;
        } else {  // zyx on disk -> xyz in memory (src/reader.cpp:385-394)
            dims = {fdims[2], fdims[1], fdims[0]};
            ms.resize(raw.size());
            const int nx = dims[0], ny = dims[1], nz = dims[2];
            for (int z = 0; z < nz; ++z)
                for (int y = 0; y < ny; ++y)
                    for (int x = 0; x < nx; ++x) ms[((size_t)x * ny + y) * nz + z] = raw[((size_t)z * ny + y) * nx + x];
        }
        finish_ms();
    }

    std::vector<double> volume_fractions;

    // NumPy .npy (format 1.0-3.0), C-ordered 3-D array of uint8 / uint16: the same image an HDF5 dataset would hold, for hosts
    // without libhdf5.  Axis order on disk follows the HDF5 convention ("zyx" unless microstructure.permute_order says "xyz").
    static bool npy_read(const std::string &fn, std::vector<int> &dims, std::vector<uint16_t> &data, std::string &err)
    {
        std::ifstream in(fn, std::ios::binary);
        if (!in) return err = "cannot open file", false;
        char magic[8];
        in.read(magic, 8);
        if (!in || std::string(magic, 6) != "\x93NUMPY") return err = "not a .npy file", false;
        size_t hlen = 0;
        unsigned char lb[4] = {0, 0, 0, 0};
        in.read((char *)lb, magic[6] >= 2 ? 4 : 2);
        hlen = lb[0] | (lb[1] << 8) | ((size_t)lb[2] << 16) | ((size_t)lb[3] << 24);
        std::string hdr(hlen, ' ');
        in.read(&hdr[0], (std::streamsize)hlen);
        int esz = 0;
        if (hdr.find("'|u1'") != std::string::npos) esz = 1;
        else if (hdr.find("'<u2'") != std::string::npos) esz = 2;
        else return err = "dtype must be uint8 or little-endian uint16", false;
        if (hdr.find("'fortran_order': False") == std::string::npos) return err = "array must be C-ordered", false;
        const size_t sp = hdr.find("'shape': (");
        if (sp == std::string::npos) return err = "no shape in header", false;
        dims.clear();
        for (size_t i = sp + 10; i < hdr.size() && hdr[i] != ')';) {
            if (isdigit((unsigned char)hdr[i])) {
                size_t j = i;
                while (j < hdr.size() && isdigit((unsigned char)hdr[j])) ++j;
                dims.push_back(std::stoi(hdr.substr(i, j - i)));
                i = j;
            } else {
                ++i;
            }
        }
        if (dims.size() != 3) return err = "array must be 3-dimensional", false;
        const size_t n = (size_t)dims[0] * dims[1] * dims[2];
        std::vector<unsigned char> raw(n * esz);
        in.read((char *)raw.data(), (std::streamsize)raw.size());
        if (!in) return err = "file is shorter than its header says", false;
        data.resize(n);
        for (size_t i = 0; i < n; ++i) data[i] = esz == 1 ? raw[i] : (uint16_t)(raw[2 * i] | (raw[2 * i + 1] << 8));
        return true;
    }

  private:
    void finish_ms()
    {
        if (L.size() != 3) throw std::runtime_error("microstructure.L must have 3 entries");
        l_e = {L[0] / dims[0], L[1] / dims[1], L[2] / dims[2]};
        if (dims[0] / 4 < world_size) throw std::runtime_error("[ERROR] Number of processes * 4 must be <= n_x");
        if (dims[0] % world_size || dims[1] % world_size) throw std::runtime_error("n_x and n_y must be divisible by the number of ranks");
        // ComputeVolumeFractions (src/reader.cpp:13-61)
        uint16_t mx = 0, mn = 65535;
        for (uint16_t v : ms) mx = std::max(mx, v), mn = std::min(mn, v);
        n_mat = (int)mx - (int)mn + 1;
        volume_fractions.assign(n_mat, 0.0);
        for (uint16_t v : ms) volume_fractions[v - mn] += 1.0;
        for (double &v : volume_fractions) v /= (double)ms.size();
        // slab of this rank (src/reader.cpp:311-331: fftw_mpi_local_size_many_transposed hands out n_x/P x-planes and n_y/P y-rows);
        // every rank has read the whole image, the volume fractions above are global like the reference's (reader.cpp:40-57)
        local_n0 = dims[0] / world_size, local_0_start = world_rank * local_n0;
        local_n1 = dims[1] / world_size, local_1_start = world_rank * local_n1;
        if (world_size > 1) {
            const size_t plane = (size_t)dims[1] * dims[2];
            std::vector<uint16_t> slab(ms.begin() + (size_t)local_0_start * plane, ms.begin() + (size_t)(local_0_start + local_n0) * plane);
            ms.swap(slab);
        }
    }
};

}  // namespace fans
