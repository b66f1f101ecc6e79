// h5mini.hpp — minimal, dependency-free (zlib only) reader for the HDF5 microstructure files FANS consumes
// (src/reader.cpp:227-411 reads them with the parallel HDF5 library, which this image does not have).
// Supported subset — what h5py / MSUtils write for an image dataset:
//   superblock v0/v1, old-style groups (symbol table: v1 B-tree + local heap) and new-style compact groups (link messages),
//   v1 object headers incl. continuation blocks,
//   dataspace v1/v2, fixed-point datatypes of 1/2/4/8 bytes (little endian), data layout v3 contiguous or chunked
//   (v1 chunk B-tree, any number of chunks), optional deflate filter, string attribute `permute_order`;
//   compact (in-header) attributes of a dataset, message versions 1-3: integer scalars and fixed- or variable-length strings
//   (h5mini_read_attributes: what GBDiffusion reads — num_crystals, num_GB, GBVoxelInfo; GBDiffusion.h:51-58).
// Anything else is reported as an error string, never guessed.
#pragma once
#include <zlib.h>

#include <cstdint>
#include <cstring>
#include <fstream>
#include <map>
#include <string>
#include <vector>

namespace fans {
namespace h5mini {

struct File {
    std::vector<unsigned char> b;
    int so = 8, sl = 8;  // size of offsets / lengths
    uint64_t rd(size_t p, int n) const
    {
        uint64_t v = 0;
        for (int i = 0; i < n; ++i) v |= (uint64_t)b.at(p + i) << (8 * i);
        return v;
    }
};

struct Msg {
    int type;
    size_t pos, size;
};

// all messages of a v1 object header at `addr` (following continuation messages)
static inline bool object_messages(const File &f, uint64_t addr, std::vector<Msg> &out, std::string &err)
{
    if (f.b.at(addr) != 1) {
        err = "only version-1 object headers are supported";
        return false;
    }
    const int nmsg = (int)f.rd(addr + 2, 2);
    const uint64_t hsize = f.rd(addr + 8, 4);
    std::vector<std::pair<size_t, size_t>> blocks{{(size_t)addr + 16, (size_t)hsize}};
    int seen = 0;
    for (size_t bi = 0; bi < blocks.size() && seen < nmsg; ++bi) {
        size_t p = blocks[bi].first;
        const size_t end = p + blocks[bi].second;
        while (p + 8 <= end && seen < nmsg) {
            const int type = (int)f.rd(p, 2);
            const size_t sz = (size_t)f.rd(p + 2, 2);
            out.push_back({type, p + 8, sz});
            if (type == 0x10) blocks.push_back({(size_t)f.rd(p + 8, f.so), (size_t)f.rd(p + 8 + f.so, f.sl)});
            p += 8 + sz;
            ++seen;
        }
    }
    return true;
}

// resolve `name` inside the old-style group whose object header is at `addr`
static inline bool group_lookup(const File &f, uint64_t addr, const std::string &name, uint64_t &child, std::string &err)
{
    std::vector<Msg> msgs;
    if (!object_messages(f, addr, msgs, err)) return false;
    for (const Msg &m : msgs) {
        if (m.type != 0x11) continue;  // symbol table message: B-tree address, local heap address
        const uint64_t btree = f.rd(m.pos, f.so), heap = f.rd(m.pos + f.so, f.so);
        if (std::memcmp(&f.b.at(heap), "HEAP", 4) != 0) {
            err = "bad local heap signature";
            return false;
        }
        const uint64_t heap_data = f.rd(heap + 8 + 2 * f.sl, f.so);
        std::vector<uint64_t> nodes{btree};
        while (!nodes.empty()) {
            const uint64_t n = nodes.back();
            nodes.pop_back();
            if (std::memcmp(&f.b.at(n), "TREE", 4) == 0) {
                const int used = (int)f.rd(n + 6, 2);
                size_t p = n + 8 + 2 * f.so;  // keys/children: key0 child0 key1 ...
                for (int i = 0; i < used; ++i) {
                    p += f.sl;  // key
                    nodes.push_back(f.rd(p, f.so));
                    p += f.so;
                }
            } else if (std::memcmp(&f.b.at(n), "SNOD", 4) == 0) {
                const int nsym = (int)f.rd(n + 6, 2);
                size_t p = n + 8;
                for (int i = 0; i < nsym; ++i) {
                    const uint64_t name_off = f.rd(p, f.so), ohdr = f.rd(p + f.so, f.so);
                    const char *s = (const char *)&f.b.at(heap_data + name_off);
                    if (name == s) {
                        child = ohdr;
                        return true;
                    }
                    p += 2 * f.so + 4 + 4 + 16;
                }
            } else {
                err = "unexpected node in group B-tree";
                return false;
            }
        }
        err = "'" + name + "' not found";
        return false;
    }
    // new-style "compact" group (libver >= 1.8 writes these for small groups): hard links stored as Link messages (0x06)
    bool any_link = false;
    for (const Msg &m : msgs) {
        if (m.type != 0x06) continue;
        any_link = true;
        size_t p = m.pos;
        if (f.b.at(p) != 1) continue;
        const int flags = f.b.at(p + 1);
        p += 2;
        int ltype = 0;
        if (flags & 0x08) ltype = f.b.at(p++);
        if (flags & 0x04) p += 8;  // creation order
        if (flags & 0x10) p += 1;  // character set
        const int lsz = 1 << (flags & 3);
        const size_t nlen = (size_t)f.rd(p, lsz);
        p += lsz;
        const std::string nm((const char *)&f.b.at(p), nlen);
        p += nlen;
        if (nm == name) {
            if (ltype != 0) {
                err = "'" + name + "' is not a hard link";
                return false;
            }
            child = f.rd(p, f.so);
            return true;
        }
    }
    err = any_link ? "'" + name + "' not found" : "group uses neither a symbol table nor compact link messages (dense link storage is not supported)";
    return false;
}

static inline bool inflate_chunk(const unsigned char *src, size_t n, std::vector<unsigned char> &dst, std::string &err)
{
    uLongf len = dst.size();
    const int rc = uncompress(dst.data(), &len, src, n);
    if (rc != Z_OK || len != dst.size()) {
        err = "deflate chunk did not inflate to the chunk size";
        return false;
    }
    return true;
}

}  // namespace h5mini

// Reads an integer dataset of rank 3 into uint16 (the cast of src/reader.cpp:260, H5T_NATIVE_USHORT).  dims = on-disk dims.
inline bool h5mini_read_dataset(const std::string &file, const std::string &dataset, std::vector<int> &dims, std::vector<uint16_t> &data,
                                std::string &permute_order, std::string &err)
{
    using namespace h5mini;
    File f;
    {
        std::ifstream in(file, std::ios::binary);
        if (!in) {
            err = "cannot open file";
            return false;
        }
        f.b.assign(std::istreambuf_iterator<char>(in), std::istreambuf_iterator<char>());
    }
    try {
        static const unsigned char sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
        if (f.b.size() < 96 || std::memcmp(f.b.data(), sig, 8) != 0) {
            err = "not an HDF5 file";
            return false;
        }
        const int sbv = f.b[8];
        if (sbv > 1) {
            err = "superblock version " + std::to_string(sbv) + " not supported (only 0/1)";
            return false;
        }
        f.so = f.b[13];
        f.sl = f.b[14];
        size_t p = 24 + (sbv == 1 ? 4 : 0);
        p += 4 * f.so;                        // base, free-space, end-of-file, driver info
        uint64_t obj = f.rd(p + f.so, f.so);  // root symbol table entry: link name offset, object header address
        // walk the path
        size_t s = 0;
        while (s < dataset.size()) {
            while (s < dataset.size() && dataset[s] == '/') ++s;
            size_t e = dataset.find('/', s);
            if (e == std::string::npos) e = dataset.size();
            if (e > s) {
                uint64_t child = 0;
                if (!group_lookup(f, obj, dataset.substr(s, e - s), child, err)) return false;
                obj = child;
            }
            s = e;
        }
        std::vector<Msg> msgs;
        if (!object_messages(f, obj, msgs, err)) return false;
        int elem = 0, rank = 0, layout_class = -1;
        std::vector<uint64_t> dsz, chunk;
        uint64_t data_addr = 0, btree = 0;
        bool deflate = false;
        permute_order = "zyx";  // default when the attribute is absent (include/reader.h:54)
        for (const Msg &m : msgs) {
            if (m.type == 0x01) {  // dataspace
                const int v = f.b.at(m.pos);
                rank = f.b.at(m.pos + 1);
                const size_t q = m.pos + (v == 1 ? 8 : 4);
                for (int i = 0; i < rank; ++i) dsz.push_back(f.rd(q + (size_t)i * f.sl, f.sl));
            } else if (m.type == 0x03) {  // datatype
                const int cls = f.b.at(m.pos) & 0x0f;
                if (cls != 0) {
                    err = "dataset is not of an integer type";
                    return false;
                }
                if (f.b.at(m.pos + 1) & 1) {
                    err = "big-endian data not supported";
                    return false;
                }
                elem = (int)f.rd(m.pos + 4, 4);
            } else if (m.type == 0x08) {  // layout
                if (f.b.at(m.pos) != 3) {
                    err = "only data layout message version 3 is supported";
                    return false;
                }
                layout_class = f.b.at(m.pos + 1);
                if (layout_class == 1) {
                    data_addr = f.rd(m.pos + 2, f.so);
                } else if (layout_class == 2) {
                    const int nd = f.b.at(m.pos + 2);
                    btree = f.rd(m.pos + 3, f.so);
                    for (int i = 0; i < nd; ++i) chunk.push_back(f.rd(m.pos + 3 + f.so + 4 * (size_t)i, 4));
                } else {
                    err = "compact layout not supported";
                    return false;
                }
            } else if (m.type == 0x0b) {  // filter pipeline
                const int v = f.b.at(m.pos), nf = f.b.at(m.pos + 1);
                size_t q = m.pos + (v == 1 ? 8 : 2);
                for (int i = 0; i < nf; ++i) {
                    const int id = (int)f.rd(q, 2);
                    const int name_len = (v == 1 || id >= 256) ? (int)f.rd(q + 2, 2) : 0;
                    const int ncv = (int)f.rd(q + (v == 1 || id >= 256 ? 6 : 4), 2);
                    if (id != 1) {
                        err = "unsupported filter id " + std::to_string(id) + " (only deflate)";
                        return false;
                    }
                    deflate = true;
                    q += (v == 1 || id >= 256 ? 8 : 6) + (size_t)((name_len + 7) / 8 * 8) + 4 * (size_t)ncv;
                    if (v == 1 && (ncv & 1)) q += 4;
                }
            } else if (m.type == 0x0c) {  // attribute (v1): name size, datatype size, dataspace size, name, ...
                const int v = f.b.at(m.pos);
                if (v != 1) continue;
                const size_t nsz = (size_t)f.rd(m.pos + 2, 2), tsz = (size_t)f.rd(m.pos + 4, 2), ssz = (size_t)f.rd(m.pos + 6, 2);
                const std::string an((const char *)&f.b.at(m.pos + 8));
                if (an != "permute_order") continue;
                const size_t pad = 8;
                const size_t q = m.pos + 8 + (nsz + pad - 1) / pad * pad + (tsz + pad - 1) / pad * pad + (ssz + pad - 1) / pad * pad;
                std::string val;
                for (size_t i = q; i < m.pos + m.size && f.b.at(i) != 0; ++i) val.push_back((char)f.b.at(i));
                if (val == "xyz" || val == "zyx") permute_order = val;
            }
        }
        // a results file stores scalar fields as [Z][Y][X][1] (include/solver.h:667-668): trailing dimensions of extent 1 are dropped
        while (rank > 3 && dsz[rank - 1] == 1) --rank;
        if (rank != 3 || elem < 1 || elem > 8 || layout_class < 0) {
            err = "dataset must be a rank-3 integer array";
            return false;
        }
        dims = {(int)dsz[0], (int)dsz[1], (int)dsz[2]};
        const size_t total = (size_t)dsz[0] * dsz[1] * dsz[2];
        std::vector<unsigned char> raw(total * elem);
        if (layout_class == 1) {
            std::memcpy(raw.data(), &f.b.at(data_addr), raw.size());
        } else {
            if (chunk.size() != 4 || (int)chunk[3] != elem) {
                err = "unexpected chunk rank";
                return false;
            }
            const size_t cbytes = (size_t)chunk[0] * chunk[1] * chunk[2] * elem;
            std::vector<unsigned char> cbuf(cbytes);
            std::vector<uint64_t> nodes{btree};
            while (!nodes.empty()) {
                const uint64_t n = nodes.back();
                nodes.pop_back();
                if (std::memcmp(&f.b.at(n), "TREE", 4) != 0 || f.b.at(n + 4) != 1) {
                    err = "bad chunk B-tree node";
                    return false;
                }
                const int level = f.b.at(n + 5), used = (int)f.rd(n + 6, 2);
                size_t q = n + 8 + 2 * f.so;
                const size_t keysz = 8 + 8 * 4;  // chunk size, filter mask, (rank+1) offsets of 8 bytes
                for (int i = 0; i < used; ++i) {
                    const uint64_t csize = f.rd(q, 4);
                    uint64_t off[3] = {f.rd(q + 8, 8), f.rd(q + 16, 8), f.rd(q + 24, 8)};
                    const uint64_t child = f.rd(q + keysz, f.so);
                    q += keysz + f.so;
                    if (level > 0) {
                        nodes.push_back(child);
                        continue;
                    }
                    if (deflate) {
                        if (!inflate_chunk(&f.b.at(child), csize, cbuf, err)) return false;
                    } else {
                        std::memcpy(cbuf.data(), &f.b.at(child), cbytes);
                    }
                    for (uint64_t a = 0; a < chunk[0] && off[0] + a < dsz[0]; ++a)
                        for (uint64_t bq = 0; bq < chunk[1] && off[1] + bq < dsz[1]; ++bq) {
                            const uint64_t nrow = std::min<uint64_t>(chunk[2], dsz[2] - off[2]);
                            std::memcpy(&raw[(((off[0] + a) * dsz[1] + off[1] + bq) * dsz[2] + off[2]) * elem],
                                        &cbuf[((a * chunk[1] + bq) * chunk[2]) * elem], nrow * elem);
                        }
                }
            }
        }
        data.resize(total);
        for (size_t i = 0; i < total; ++i) {
            uint64_t v = 0;
            for (int k = 0; k < elem; ++k) v |= (uint64_t)raw[i * elem + k] << (8 * k);
            data[i] = (uint16_t)v;
        }
        return true;
    } catch (const std::out_of_range &) {
        err = "truncated or malformed file";
        return false;
    }
}

// Attributes stored in the object header of `dataset`: integers as decimal text in `ints`, strings in `strs`.  Attributes in dense
// storage (fractal heap) or of other types are skipped; `err` is only set when the file / dataset itself cannot be read.
struct H5Attrs {
    std::map<std::string, long long> ints;
    std::map<std::string, std::string> strs;
};

inline bool h5mini_read_attributes(const std::string &file, const std::string &dataset, H5Attrs &out, std::string &err)
{
    using namespace h5mini;
    File f;
    {
        std::ifstream in(file, std::ios::binary);
        if (!in) {
            err = "cannot open file";
            return false;
        }
        f.b.assign(std::istreambuf_iterator<char>(in), std::istreambuf_iterator<char>());
    }
    try {
        static const unsigned char sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
        if (f.b.size() < 96 || std::memcmp(f.b.data(), sig, 8) != 0 || f.b[8] > 1) {
            err = "not an HDF5 file with a version 0/1 superblock";
            return false;
        }
        f.so = f.b[13];
        f.sl = f.b[14];
        size_t p = 24 + (f.b[8] == 1 ? 4 : 0) + 4 * (size_t)f.so;
        uint64_t obj = f.rd(p + f.so, f.so);
        size_t s = 0;
        while (s < dataset.size()) {
            while (s < dataset.size() && dataset[s] == '/') ++s;
            size_t e = dataset.find('/', s);
            if (e == std::string::npos) e = dataset.size();
            if (e > s) {
                uint64_t child = 0;
                if (!group_lookup(f, obj, dataset.substr(s, e - s), child, err)) return false;
                obj = child;
            }
            s = e;
        }
        std::vector<Msg> msgs;
        if (!object_messages(f, obj, msgs, err)) return false;
        for (const Msg &m : msgs) {
            if (m.type != 0x0c) continue;
            const int v = f.b.at(m.pos);
            if (v < 1 || v > 3) continue;
            const size_t nsz = (size_t)f.rd(m.pos + 2, 2), tsz = (size_t)f.rd(m.pos + 4, 2), ssz = (size_t)f.rd(m.pos + 6, 2);
            size_t q = m.pos + 8 + (v == 3 ? 1 : 0);   // v3: + name character-set encoding
            auto adv = [&](size_t n) { return v == 1 ? (n + 7) / 8 * 8 : n; };   // v1 pads every part to 8 bytes
            const std::string name((const char *)&f.b.at(q));
            q += adv(nsz);
            const size_t tpos = q;
            q += adv(tsz);
            const size_t spos = q;
            q += adv(ssz);
            const int cls = f.b.at(tpos) & 0x0f;
            const size_t tsize = (size_t)f.rd(tpos + 4, 4);
            const int srank = f.b.at(spos + 1);
            if (srank > 1) continue;
            if (cls == 0 && tsize >= 1 && tsize <= 8) {          // fixed-point scalar (sign-extended)
                uint64_t raw = f.rd(q, (int)tsize);
                const bool is_signed = (f.b.at(tpos + 1) & 0x08) != 0;
                long long val = (long long)raw;
                if (is_signed && tsize < 8 && (raw >> (8 * tsize - 1)) & 1) val = (long long)(raw | (~0ULL << (8 * tsize)));
                out.ints[name] = val;
            } else if (cls == 3) {                                // fixed-length string
                std::string val;
                for (size_t i = 0; i < tsize && f.b.at(q + i) != 0; ++i) val.push_back((char)f.b.at(q + i));
                out.strs[name] = val;
            } else if (cls == 9) {                                // variable length: {length, global heap collection, index}
                const uint32_t len = (uint32_t)f.rd(q, 4);
                const uint64_t gcol = f.rd(q + 4, f.so);
                const uint32_t index = (uint32_t)f.rd(q + 4 + f.so, 4);
                if (std::memcmp(&f.b.at(gcol), "GCOL", 4) != 0) continue;
                const uint64_t csize = f.rd(gcol + 8, f.sl);
                size_t o = gcol + 8 + f.sl;
                while (o + 8 + f.sl <= gcol + csize) {
                    const uint32_t oi = (uint32_t)f.rd(o, 2);
                    const uint64_t osz = f.rd(o + 8, f.sl);
                    if (oi == 0) break;   // free space object ends the collection
                    if (oi == index) {
                        out.strs[name] = std::string((const char *)&f.b.at(o + 8 + f.sl), std::min<size_t>(len, (size_t)osz));
                        break;
                    }
                    o += 8 + f.sl + (size_t)((osz + 7) / 8 * 8);
                }
            }
        }
        return true;
    } catch (const std::out_of_range &) {
        err = "truncated or malformed file";
        return false;
    }
}

}  // namespace fans
