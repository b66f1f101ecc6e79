// main.cpp — command-line front end of the GPU build, mirror of the reference's src/main.cpp:9-80:
//     FANS_gpu input.json results_dir          (reference: mpiexec -n P FANS input.json results.h5)
// Same JSON input, same load-case x time-step loop (runSolver), same console summary.  Results go through a ResultsSink;
// libhdf5 is not available in this image, so the sink writes one raw little-endian file per dataset under
//     results_dir/<dataset>_results/<prefix>/load<L>/time_step<T>/<name>.bin   + results_dir/index.jsonl (name, dtype, dims; "xyz" order)
// following the reference's HDF5 group layout (include/reader.h:331-351).
//     FANS_gpu --describe input.json [ms.u16 nx ny nz]   prints what the host derives from the input (no GPU needed).
//     FANS_gpu <input.json> results.h5 ...                writes the reference's HDF5 results layout instead (h5write.hpp).
// Several ranks (reference: mpiexec -n P, src/main.cpp:60-61; here one process per GPU, no MPI needed): start P copies with
// RANK / WORLD_SIZE / LOCAL_RANK in the environment (torchrun's names; OMPI_COMM_WORLD_* and FANS_* are accepted too) and a shared
// FANS_COMM_FILE path — rank 0 drops the 128-byte NCCL id there, the others pick it up (tools/fans_mprun.py does exactly that).
// Every rank owns n_x/P x-planes; rank r > 0 writes its field slabs to results_dir/rank<r>/ (results.rank<r>.h5), small data come
// from rank 0 only.
#include <sys/stat.h>
#include <unistd.h>

#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>

#include "h5write.hpp"
#include "solver.hpp"

using namespace fans;

struct DirSink : ResultsSink {
    std::string root, dataset;
    std::ofstream index;
    DirSink(const std::string &r, const std::string &ds) : root(r), dataset(ds)
    {
        mkdirs(root);
        index.open(root + "/index.jsonl", std::ios::app);
    }
    static void mkdirs(const std::string &p)
    {
        for (size_t i = 1; i <= p.size(); ++i)
            if (i == p.size() || p[i] == '/') mkdir(p.substr(0, i).c_str(), 0777);
    }
    void write(const std::string &name, int load_idx, int time_idx, const std::string &dtype, const std::vector<size_t> &dims, const void *data,
               bool is_field) override
    {
        const std::string dir = root + dataset + "/load" + std::to_string(load_idx) + "/time_step" + std::to_string(time_idx);
        mkdirs(dir);
        size_t n = 1;
        for (size_t d : dims) n *= d;
        const size_t esz = dtype == "f64" ? 8 : (dtype == "u16" ? 2 : 4);
        std::ofstream f(dir + "/" + name + ".bin", std::ios::binary);
        f.write((const char *)data, (std::streamsize)(n * esz));
        index << "{\"name\": \"" << name << "\", \"load\": " << load_idx << ", \"time_step\": " << time_idx << ", \"dtype\": \"" << dtype
              << "\", \"dims\": [";
        for (size_t i = 0; i < dims.size(); ++i) index << (i ? ", " : "") << dims[i];
        index << "], \"field\": " << (is_field ? "true" : "false") << ", \"order\": \"xyz\", \"path\": \"" << dir.substr(root.size()) << "/" << name
              << ".bin\"}\n";
        index.flush();
    }
};

// HDF5 results file with the reference's layout (include/reader.h:173-351): <dataset>/load<L>/time_step<T>/<name>; fields are
// transposed from the solver's [X][Y][Z][extra] to [Z][Y][X][extra] and tagged permute_order = "zyx" exactly like Reader::WriteSlab.
struct H5Sink : ResultsSink {
    h5w::Writer w;
    std::string dataset;
    H5Sink(const std::string &file, const std::string &ds) : w(file), dataset(ds) {}
    void write(const std::string &name, int load_idx, int time_idx, const std::string &dtype, const std::vector<size_t> &dims, const void *data,
               bool is_field) override
    {
        const std::string path = dataset + "/load" + std::to_string(load_idx) + "/time_step" + std::to_string(time_idx) + "/" + name;
        std::vector<uint64_t> d(dims.begin(), dims.end());
        if (!is_field || dims.size() < 3) {
            w.add_dataset(path, dtype, d, data);
            return;
        }
        const size_t X = dims[0], Y = dims[1], Z = dims[2], esz = h5w::Writer::elem_size(dtype);
        size_t extra = 1;
        for (size_t i = 3; i < dims.size(); ++i) extra *= dims[i];
        const size_t row = extra * esz;
        std::vector<unsigned char> t(X * Y * Z * row);
        const unsigned char *src = (const unsigned char *)data;
        for (size_t x = 0; x < X; ++x)
            for (size_t y = 0; y < Y; ++y)
                for (size_t z = 0; z < Z; ++z) std::memcpy(&t[((z * Y + y) * X + x) * row], src + ((x * Y + y) * Z + z) * row, row);
        d[0] = Z, d[2] = X;
        if (d.size() == 3) d.push_back(1);  // scalar fields carry an explicit extra dimension of 1 (solver.h:667-668: writeSlab(..., {1}))
        w.add_dataset(path, dtype, d, t.data(), "permute_order", "zyx");
    }
};

// writes a small file exercising every feature of h5write.hpp (nested groups, many entries per group, all element types, field
// transpose + attribute); tests/test_host_cpp.py parses it back with an independent reader.  No GPU needed.
static int h5_selftest(const char *file)
{
    H5Sink sink(file, "/img/4x3x2/ms_results/run1");
    const size_t X = 4, Y = 3, Z = 2;
    std::vector<double> f(X * Y * Z * 3);
    std::vector<uint16_t> ms(X * Y * Z);
    for (size_t x = 0; x < X; ++x)
        for (size_t y = 0; y < Y; ++y)
            for (size_t z = 0; z < Z; ++z) {
                ms[(x * Y + y) * Z + z] = (uint16_t)(100 * x + 10 * y + z);
                for (size_t c = 0; c < 3; ++c) f[((x * Y + y) * Z + z) * 3 + c] = 1000.0 * x + 100.0 * y + 10.0 * z + c + 0.5;
            }
    for (int t = 0; t < 40; ++t) {  // 40 time steps: more entries than the default leaf size of libhdf5 groups
        std::vector<double> sa(6);
        for (int i = 0; i < 6; ++i) sa[i] = t + 0.125 * i;
        sink.write("stress_average", 0, t, "f64", {6}, sa.data(), false);
    }
    sink.write("displacement", 0, 0, "f64", {X, Y, Z, 3}, f.data(), true);
    sink.write("microstructure", 0, 0, "u16", {X, Y, Z}, ms.data(), true);
    std::vector<float> ff = {1.5f, -2.25f, 3.0f};
    std::vector<int> ii = {-7, 0, 123456};
    sink.write("some_floats", 1, 0, "f32", {3}, ff.data(), false);
    sink.write("plastic_flag_like", 1, 0, "i32", {3}, ii.data(), false);
    std::vector<double> C(36);
    for (int i = 0; i < 36; ++i) C[i] = i * i;
    sink.write("homogenized_tangent", 1, 0, "f64", {6, 6}, C.data(), false);
    sink.w.close();
    return 0;
}

// writes a small grain-boundary image with the dataset attributes GBDiffusion reads (num_crystals, num_GB, GBVoxelInfo;
// GBDiffusion.h:51-69): 2 crystals (tags 0, 1) separated by two one-voxel boundary layers (tags 2, 3) with oblique normals.
// tests/test_host_cpp.py reads it back through the JSON front end (--describe).  No GPU needed.
static int gb_selftest(const char *file)
{
    const size_t X = 8, Y = 4, Z = 4;
    std::vector<uint16_t> ms(Z * Y * X);   // on disk [z][y][x]
    for (size_t z = 0; z < Z; ++z)
        for (size_t y = 0; y < Y; ++y)
            for (size_t x = 0; x < X; ++x) ms[(z * Y + y) * X + x] = (uint16_t)(x < 3 ? 0 : (x == 3 ? 2 : (x < 7 ? 1 : 3)));
    h5w::Writer w(file);
    h5w::Dataset &d = w.add_dataset("/gb/image", "u16", {Z, Y, X}, ms.data());
    d.int_attrs.push_back({"num_crystals", 2});
    d.int_attrs.push_back({"num_GB", 2});
    d.str_attrs.push_back({"GBVoxelInfo", "{\"a\": {\"GB_tag\": 2, \"GB_normal\": [1.0, 0.0, 0.0]}, "
                                          "\"b\": {\"GB_tag\": 3, \"GB_normal\": [0.6, 0.8, 0.0]}}"});
    w.close();
    return 0;
}

// ---- slab world: rank / size from the environment, NCCL id through a file (the role MPI_Init + MPI_Bcast play in the reference) ----
static int env_int(std::initializer_list<const char *> names, int dflt)
{
    for (const char *n : names)
        if (const char *v = getenv(n)) return atoi(v);
    return dflt;
}

static void slab_world(Reader &reader)
{
    reader.world_size = env_int({"FANS_WORLD_SIZE", "WORLD_SIZE", "OMPI_COMM_WORLD_SIZE"}, 1);
    reader.world_rank = env_int({"FANS_RANK", "RANK", "OMPI_COMM_WORLD_RANK"}, 0);
    reader.device = env_int({"FANS_DEVICE", "LOCAL_RANK", "OMPI_COMM_WORLD_LOCAL_RANK"}, -1);
    if (reader.world_size <= 1) {
        reader.world_size = 1, reader.world_rank = 0;
        return;
    }
    const char *f = getenv("FANS_COMM_FILE");
    if (!f) throw std::runtime_error("WORLD_SIZE > 1 needs FANS_COMM_FILE (a path all ranks share) for the NCCL id rendezvous");
    const std::string file = f;
    unsigned char id[128];
    if (reader.world_rank == 0) {
        if (fans_comm_unique_id(id) != FANS_OK) throw std::runtime_error(std::string("fans_comm_unique_id: ") + fans_last_error(nullptr));
        const std::string tmp = file + ".tmp";
        {
            std::ofstream o(tmp, std::ios::binary);
            o.write((const char *)id, 128);
        }
        if (rename(tmp.c_str(), file.c_str()) != 0) throw std::runtime_error("cannot publish " + file);
    } else {
        bool ok = false;
        for (int i = 0; i < 6000 && !ok; ++i) {   // up to 60 s
            std::ifstream in(file, std::ios::binary);
            if (in && in.read((char *)id, 128) && in.gcount() == 128) ok = true;
            else usleep(10000);
        }
        if (!ok) throw std::runtime_error("timed out waiting for the NCCL id in " + file);
    }
    if (fans_comm_create(&reader.comm, reader.world_size, reader.world_rank, id, reader.device) != FANS_OK)
        throw std::runtime_error(std::string("fans_comm_create: ") + fans_last_error(nullptr));
    if (reader.world_rank == 0) unlink(file.c_str());   // everybody holds the communicator once CommInitRank has returned
}

// field slabs of rank r > 0 go next to rank 0's; small data (averages, error history, tangent) are identical on every rank
struct RankSink : ResultsSink {
    ResultsSink &inner;
    int rank;
    RankSink(ResultsSink &s, int r) : inner(s), rank(r) {}
    void write(const std::string &name, int load_idx, int time_idx, const std::string &dtype, const std::vector<size_t> &dims, const void *data,
               bool is_field) override
    {
        if (is_field || rank == 0) inner.write(name, load_idx, time_idx, dtype, dims, data, is_field);
    }
};

static std::string rank_path(const std::string &out, int rank, bool h5)
{
    if (rank == 0) return out;
    if (h5) return out.substr(0, out.size() - 3) + ".rank" + std::to_string(rank) + ".h5";
    return out + "/rank" + std::to_string(rank);
}

static void load_raw_ms(Reader &reader, const char *file, int nx, int ny, int nz)
{
    std::ifstream in(file, std::ios::binary);
    if (!in) throw std::runtime_error(std::string("cannot open ") + file);
    std::vector<uint16_t> ms((size_t)nx * ny * nz);
    in.read((char *)ms.data(), (std::streamsize)(ms.size() * 2));
    const int d[3] = {nx, ny, nz};
    reader.SetMicrostructure(d, ms.data());
}

static int describe(Reader &reader)
{
    MaterialManager mm(reader);
    printf("{\"howmany\": %d, \"n_str\": %d, \"n_phases\": %d, \"all_linear\": %s, \"dims\": [%d, %d, %d], \"n_load_cases\": %zu,\n \"kapparef\": [",
           reader.howmany(), reader.n_str(), mm.n_phases, mm.all_linear ? "true" : "false", reader.dims[0], reader.dims[1], reader.dims[2],
           reader.load_cases.size());
    for (size_t i = 0; i < mm.kapparef_mat.a.size(); ++i) printf("%s%.17g", i ? ", " : "", mm.kapparef_mat.a[i]);
    printf("],\n \"phases\": [");
    const auto descs = mm.phase_descs();
    for (size_t p = 0; p < descs.size(); ++p) {
        printf("%s{\"model\": %d, \"local_mat\": %d, \"group_n_mat\": %d, \"params\": [", p ? ", " : "", descs[p].model, descs[p].local_mat,
               descs[p].group_n_mat);
        for (int k = 0; k < FANS_MAX_PARAMS; ++k) printf("%s%.17g", k ? ", " : "", descs[p].params[k]);
        printf("]}");
    }
    printf("],\n \"volume_fractions\": [");
    for (size_t i = 0; i < reader.volume_fractions.size(); ++i) printf("%s%.17g", i ? ", " : "", reader.volume_fractions[i]);
    printf("],\n \"mixed_M\": [");
    bool first = true;
    for (auto &lc : reader.load_cases)
        if (lc.mixed) {
            MixedBC b = lc.mbc;
            b.finalize(mm.kapparef_mat);
            printf("%s[", first ? "" : ", ");
            for (size_t i = 0; i < b.M.a.size(); ++i) printf("%s%.17g", i ? ", " : "", b.M.a[i]);
            printf("]");
            first = false;
        }
    printf("]}\n");
    return 0;
}

// runSolver, src/main.cpp:9-46
static void runSolver(Reader &reader, ResultsSink &sink)
{
    for (size_t load_path_idx = 0; load_path_idx < reader.load_cases.size(); ++load_path_idx) {
        MaterialManager *matmanager = createMaterialManager(reader);
        Solver *solver = createSolver(reader, matmanager);  // a fresh Solver (u = 0) per load case, like the reference
        for (size_t time_step_idx = 0; time_step_idx < reader.load_cases[load_path_idx].n_steps; ++time_step_idx) {
            if (reader.world_rank == 0)
                printf("\n║ ▶ Load case %zu/%zu: Time step %zu/%zu\n", load_path_idx + 1, reader.load_cases.size(), time_step_idx + 1,
                       reader.load_cases[load_path_idx].n_steps);
            if (reader.load_cases[load_path_idx].mixed) {
                solver->enableMixedBC(reader.load_cases[load_path_idx].mbc, time_step_idx);
            } else {
                solver->disableMixedBC();
                matmanager->set_gradient(reader.load_cases[load_path_idx].g0_path[time_step_idx]);
            }
            solver->solve();
            solver->postprocess(sink, (int)load_path_idx, (int)time_step_idx);
            if (reader.world_rank == 0) printf("# iterations %zu\n", solver->iter);
            if (reader.extrapolate_displacement) solver->extrapolateDisplacement();
        }
        delete solver;
        delete matmanager;
    }
}

int main(int argc, char **argv)
{
    try {
        if (argc >= 3 && std::strcmp(argv[1], "--describe") == 0) {
            Reader reader;
            reader.ReadInputFile(argv[2]);
            if (argc >= 7) load_raw_ms(reader, argv[3], atoi(argv[4]), atoi(argv[5]), atoi(argv[6]));
            else reader.ReadMS(reader.howmany());
            return describe(reader);
        }
        if (argc == 3 && std::strcmp(argv[1], "--h5selftest") == 0) return h5_selftest(argv[2]);
        if (argc == 3 && std::strcmp(argv[1], "--gbselftest") == 0) return gb_selftest(argv[2]);
        if (argc != 3 && argc != 7) {
            fprintf(stderr, "Usage: %s <input_file.json> <results_dir | results.h5> [ms.u16 nx ny nz]\n", argv[0]);
            return 10;
        }
        Reader reader;
        slab_world(reader);
        reader.ReadInputFile(argv[1]);
        if (argc == 7) load_raw_ms(reader, argv[3], atoi(argv[4]), atoi(argv[5]), atoi(argv[6]));
        else reader.ReadMS(reader.howmany());
        const std::string out = argv[2];
        const bool h5 = out.size() > 3 && out.compare(out.size() - 3, 3, ".h5") == 0;
        const std::string mine = rank_path(out, reader.world_rank, h5);
        if (h5) {  // HDF5 results file like the reference's
            H5Sink sink(mine, reader.dataset_name);
            RankSink rs(sink, reader.world_rank);
            runSolver(reader, rs);
            sink.w.close();
        } else {  // directory of raw arrays + index.jsonl
            DirSink sink(mine, reader.dataset_name);
            RankSink rs(sink, reader.world_rank);
            runSolver(reader, rs);
        }
        if (reader.comm) fans_comm_destroy(reader.comm);
        return 0;
    } catch (const std::exception &e) {
        fprintf(stderr, "ERROR: %s\n", e.what());
        return 10;
    }
}
