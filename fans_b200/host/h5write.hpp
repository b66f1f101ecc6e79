// h5write.hpp — minimal, dependency-free HDF5 writer for the results file FANS produces (include/reader.h:173-351 writes it with the
// parallel HDF5 library, which this image does not have).  It emits the most conservative on-disk subset of the HDF5 File Format
// Specification (what libhdf5 itself writes with H5F_LIBVER_EARLIEST, readable by every libhdf5 / h5py version):
//   superblock version 0 (8-byte offsets and lengths),
//   old-style groups: version-1 object header with one Symbol Table message -> version-1 B-tree ("TREE", one leaf) -> one symbol
//   table node ("SNOD", entries sorted by name) + local heap ("HEAP") holding the link names,
//   datasets: version-1 object header with Dataspace (v1, simple, no max dims), Datatype (v1: IEEE f64/f32 little endian, u16, i32),
//   Fill Value (v2, default), Data Layout (v3, contiguous) and optionally one scalar fixed-length string Attribute (v1),
//   raw data stored contiguously, little endian.
// The symbol-table leaf size K is chosen at close() from the largest group (stored in the superblock, which is where libhdf5 reads
// it from), so every group needs exactly one leaf.  Raw data is appended as it arrives; all metadata and the superblock are written
// by close() — a file that was not closed is not a valid HDF5 file.
// Layout written by H5Sink below = the reference's: <dataset>_results/<prefix>/load<L>/time_step<T>/<field>; fields as
// [Z][Y][X][extra] with the attribute permute_order = "zyx" (reader.h:221-252), small data with the dims the caller gives.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace fans {
namespace h5w {

static const uint64_t UNDEF = 0xffffffffffffffffull;

struct Buf {
    std::vector<unsigned char> b;
    void u8(unsigned v) { b.push_back((unsigned char)v); }
    void le(uint64_t v, int n)
    {
        for (int i = 0; i < n; ++i) b.push_back((unsigned char)(v >> (8 * i)));
    }
    void bytes(const void *p, size_t n) { b.insert(b.end(), (const unsigned char *)p, (const unsigned char *)p + n); }
    void zeros(size_t n) { b.insert(b.end(), n, 0); }
    void pad8()
    {
        while (b.size() % 8) b.push_back(0);
    }
};

struct Dataset {
    std::string dtype;  // "f64", "f32", "u16", "i32"
    std::vector<uint64_t> dims;
    uint64_t addr = 0, nbytes = 0;
    std::string attr_name, attr_value;  // optional scalar string attribute (value stored with its terminating NUL)
    std::vector<std::pair<std::string, std::string>> str_attrs;   // further scalar string attributes
    std::vector<std::pair<std::string, long long>> int_attrs;     // scalar 64-bit integer attributes (e.g. num_crystals, GBDiffusion.h:54-55)
};

struct Group {
    std::map<std::string, Group> groups;   // std::map keeps the names in strcmp order, the order a symbol table node needs
    std::map<std::string, Dataset> dsets;
    uint64_t ohdr = 0, btree = 0, heap = 0;
    size_t entries() const { return groups.size() + dsets.size(); }
};

class Writer {
  public:
    explicit Writer(const std::string &path) : path_(path)
    {
        f_ = std::fopen(path.c_str(), "wb");
        if (!f_) throw std::runtime_error("h5write: cannot create " + path);
        std::vector<unsigned char> z(DATA_START, 0);  // room for the superblock, written by close()
        put(z.data(), z.size());
    }
    ~Writer()
    {
        if (f_) {
            try {
                close();
            } catch (...) {
            }
        }
    }
    static size_t elem_size(const std::string &dt)
    {
        if (dt == "f64") return 8;
        if (dt == "f32" || dt == "i32") return 4;
        if (dt == "u16") return 2;
        throw std::runtime_error("h5write: unsupported element type " + dt);
    }
    // path = "/a/b/name"; raw data is appended to the file now
    Dataset &add_dataset(const std::string &path, const std::string &dtype, const std::vector<uint64_t> &dims, const void *data,
                         const std::string &attr_name = "", const std::string &attr_value = "")
    {
        if (!f_) throw std::runtime_error("h5write: file already closed");
        std::vector<std::string> parts;
        size_t i = 0;
        while (i < path.size()) {
            while (i < path.size() && path[i] == '/') ++i;
            size_t j = i;
            while (j < path.size() && path[j] != '/') ++j;
            if (j > i) parts.push_back(path.substr(i, j - i));
            i = j;
        }
        if (parts.empty()) throw std::runtime_error("h5write: empty dataset path");
        Group *g = &root_;
        for (size_t k = 0; k + 1 < parts.size(); ++k) {
            if (g->dsets.count(parts[k])) throw std::runtime_error("h5write: '" + parts[k] + "' is a dataset, not a group");
            g = &g->groups[parts[k]];
        }
        const std::string &name = parts.back();
        if (g->dsets.count(name) || g->groups.count(name)) throw std::runtime_error("h5write: '" + path + "' exists already");
        Dataset d;
        d.dtype = dtype;
        d.dims = dims;
        uint64_t n = 1;
        for (uint64_t v : dims) n *= v;
        d.nbytes = n * elem_size(dtype);
        d.attr_name = attr_name;
        d.attr_value = attr_value;
        align8();
        d.addr = pos_;
        put(data, d.nbytes);
        g->dsets[name] = d;
        return g->dsets[name];   // std::map nodes are stable: the caller may add attributes until close()
    }
    void close()
    {
        if (!f_) return;
        size_t mx = 1;
        max_entries(root_, mx);
        leaf_k_ = (int)((mx + 1) / 2);
        if (leaf_k_ < 4) leaf_k_ = 4;
        if (leaf_k_ > 65535) throw std::runtime_error("h5write: a group has more than 131070 entries");
        emit_group(root_);
        align8();
        const uint64_t eof = pos_;
        Buf sb;
        const unsigned char sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
        sb.bytes(sig, 8);
        sb.u8(0);  // superblock version
        sb.u8(0);  // free-space storage version
        sb.u8(0);  // root group symbol table entry version
        sb.u8(0);
        sb.u8(0);  // shared header message format version
        sb.u8(8);  // size of offsets
        sb.u8(8);  // size of lengths
        sb.u8(0);
        sb.le((uint64_t)leaf_k_, 2);   // group leaf node K
        sb.le(INTERNAL_K, 2);          // group internal node K
        sb.le(0, 4);                   // file consistency flags
        sb.le(0, 8);                   // base address
        sb.le(UNDEF, 8);               // free-space info
        sb.le(eof, 8);                 // end of file address
        sb.le(UNDEF, 8);               // driver information block
        symbol_entry(sb, 0, root_.ohdr, true, root_.btree, root_.heap);
        std::fseek(f_, 0, SEEK_SET);
        if (std::fwrite(sb.b.data(), 1, sb.b.size(), f_) != sb.b.size()) throw std::runtime_error("h5write: write failed");
        std::fclose(f_);
        f_ = nullptr;
    }

  private:
    static const uint64_t DATA_START = 2048;
    static const uint64_t INTERNAL_K = 16;
    std::string path_;
    FILE *f_ = nullptr;
    uint64_t pos_ = 0;
    Group root_;
    int leaf_k_ = 4;

    void put(const void *p, size_t n)
    {
        if (n && std::fwrite(p, 1, n, f_) != n) throw std::runtime_error("h5write: write failed (disk full?)");
        pos_ += n;
    }
    void align8()
    {
        static const unsigned char z[8] = {0};
        if (pos_ % 8) put(z, 8 - pos_ % 8);
    }
    uint64_t put_block(const Buf &b)
    {
        align8();
        const uint64_t at = pos_;
        put(b.b.data(), b.b.size());
        return at;
    }
    static void max_entries(const Group &g, size_t &mx)
    {
        if (g.entries() > mx) mx = g.entries();
        for (const auto &kv : g.groups) max_entries(kv.second, mx);
    }
    // symbol table entry (40 bytes): link name offset, object header address, cache type, reserved, scratch pad
    static void symbol_entry(Buf &b, uint64_t name_off, uint64_t ohdr, bool is_group, uint64_t btree, uint64_t heap)
    {
        b.le(name_off, 8);
        b.le(ohdr, 8);
        b.le(is_group ? 1 : 0, 4);
        b.le(0, 4);
        if (is_group) {
            b.le(btree, 8);
            b.le(heap, 8);
        } else {
            b.zeros(16);
        }
    }
    static void msg_header(Buf &b, unsigned type, size_t size, unsigned flags)
    {
        b.le(type, 2);
        b.le(size, 2);
        b.u8(flags);
        b.zeros(3);
    }
    static void datatype_body(Buf &b, const std::string &dt)  // padded to a multiple of 8
    {
        if (dt == "f64" || dt == "f32") {
            const bool d = dt == "f64";
            b.u8(0x11);               // version 1, class 1 (floating point)
            b.u8(0x20);               // little endian, no padding, mantissa normalisation 2 (msb implied)
            b.u8(d ? 63 : 31);        // sign bit location
            b.u8(0);
            b.le(d ? 8 : 4, 4);       // size
            b.le(0, 2);               // bit offset
            b.le(d ? 64 : 32, 2);     // precision
            b.u8(d ? 52 : 23);        // exponent location
            b.u8(d ? 11 : 8);         // exponent size
            b.u8(0);                  // mantissa location
            b.u8(d ? 52 : 23);        // mantissa size
            b.le(d ? 1023 : 127, 4);  // exponent bias
        } else {
            const bool s = dt == "i32";
            const unsigned sz = s ? 4 : 2;
            b.u8(0x10);               // version 1, class 0 (fixed point)
            b.u8(s ? 0x08 : 0x00);    // little endian, zero padding, signed flag
            b.u8(0);
            b.u8(0);
            b.le(sz, 4);
            b.le(0, 2);               // bit offset
            b.le(8 * sz, 2);          // precision
        }
        b.pad8();
    }
    uint64_t emit_dataset(const Dataset &d)
    {
        Buf m;  // the messages
        int nmsg = 0;
        {   // Dataspace, version 1
            Buf x;
            x.u8(1);
            x.u8((unsigned)d.dims.size());
            x.u8(0);
            x.zeros(5);
            for (uint64_t v : d.dims) x.le(v, 8);
            msg_header(m, 0x0001, x.b.size(), 0);
            m.bytes(x.b.data(), x.b.size());
            ++nmsg;
        }
        {   // Datatype
            Buf x;
            datatype_body(x, d.dtype);
            msg_header(m, 0x0003, x.b.size(), 1);
            m.bytes(x.b.data(), x.b.size());
            ++nmsg;
        }
        {   // Fill Value, version 2: allocation time late, write time "if set", default fill value (defined, size 0)
            Buf x;
            x.u8(2);
            x.u8(2);
            x.u8(2);
            x.u8(1);
            x.le(0, 4);
            msg_header(m, 0x0005, x.b.size(), 1);
            m.bytes(x.b.data(), x.b.size());
            ++nmsg;
        }
        {   // Data Layout, version 3, contiguous
            Buf x;
            x.u8(3);
            x.u8(1);
            x.le(d.nbytes ? d.addr : UNDEF, 8);
            x.le(d.nbytes, 8);
            x.pad8();
            msg_header(m, 0x0008, x.b.size(), 0);
            m.bytes(x.b.data(), x.b.size());
            ++nmsg;
        }
        if (!d.attr_name.empty()) {  // Attribute, version 1: scalar fixed-length (null-terminated, ASCII) string
            Buf x;
            const size_t nlen = d.attr_name.size() + 1, vlen = d.attr_value.size() + 1;
            x.u8(1);
            x.u8(0);
            x.le(nlen, 2);
            x.le(8, 2);  // datatype message size
            x.le(8, 2);  // dataspace message size
            x.bytes(d.attr_name.c_str(), nlen);
            x.pad8();
            x.u8(0x13);  // version 1, class 3 (string); null terminated, ASCII
            x.zeros(3);
            x.le(vlen, 4);
            x.u8(1);     // dataspace version 1, rank 0 (scalar)
            x.zeros(7);
            x.bytes(d.attr_value.c_str(), vlen);
            x.pad8();
            msg_header(m, 0x000c, x.b.size(), 0);
            m.bytes(x.b.data(), x.b.size());
            ++nmsg;
        }
        for (const auto &kv : d.str_attrs) {   // more scalar fixed-length strings, same encoding as above
            Buf x;
            const size_t nlen = kv.first.size() + 1, vlen = kv.second.size() + 1;
            if (vlen > 60000) throw std::runtime_error("h5write: attribute '" + kv.first + "' does not fit an object header message");
            x.u8(1);
            x.u8(0);
            x.le(nlen, 2);
            x.le(8, 2);
            x.le(8, 2);
            x.bytes(kv.first.c_str(), nlen);
            x.pad8();
            x.u8(0x13);
            x.zeros(3);
            x.le(vlen, 4);
            x.u8(1);
            x.zeros(7);
            x.bytes(kv.second.c_str(), vlen);
            x.pad8();
            msg_header(m, 0x000c, x.b.size(), 0);
            m.bytes(x.b.data(), x.b.size());
            ++nmsg;
        }
        for (const auto &kv : d.int_attrs) {   // scalar little-endian signed 64-bit integers
            Buf x;
            const size_t nlen = kv.first.size() + 1;
            x.u8(1);
            x.u8(0);
            x.le(nlen, 2);
            x.le(12, 2);  // datatype message: 8 bytes + 4 bytes of fixed-point properties
            x.le(8, 2);
            x.bytes(kv.first.c_str(), nlen);
            x.pad8();
            x.u8(0x10);   // version 1, class 0 (fixed point)
            x.u8(0x08);   // little endian, signed
            x.zeros(2);
            x.le(8, 4);   // size
            x.le(0, 2);   // bit offset
            x.le(64, 2);  // precision
            x.pad8();
            x.u8(1);      // dataspace version 1, rank 0
            x.zeros(7);
            x.le((uint64_t)kv.second, 8);
            msg_header(m, 0x000c, x.b.size(), 0);
            m.bytes(x.b.data(), x.b.size());
            ++nmsg;
        }
        Buf h;
        h.u8(1);
        h.u8(0);
        h.le((uint64_t)nmsg, 2);
        h.le(1, 4);             // object reference count
        h.le(m.b.size(), 4);    // object header size
        h.zeros(4);             // pad to 8
        h.bytes(m.b.data(), m.b.size());
        return put_block(h);
    }
    void emit_group(Group &g)
    {
        // children first (their object header addresses go into this group's symbol table node)
        std::map<std::string, uint64_t> ds_addr;
        for (auto &kv : g.dsets) ds_addr[kv.first] = emit_dataset(kv.second);
        for (auto &kv : g.groups) emit_group(kv.second);
        // names in strcmp order across both kinds
        std::map<std::string, int> names;  // 0 dataset, 1 group
        for (auto &kv : g.dsets) names[kv.first] = 0;
        for (auto &kv : g.groups) names[kv.first] = 1;
        // local heap: offset 0 holds the empty string, then the NUL-terminated names, each on an 8-byte boundary
        Buf hd;
        hd.zeros(8);
        std::map<std::string, uint64_t> off;
        for (auto &kv : names) {
            off[kv.first] = hd.b.size();
            hd.bytes(kv.first.c_str(), kv.first.size() + 1);
            hd.pad8();
        }
        Buf hp;
        hp.bytes("HEAP", 4);
        hp.u8(0);
        hp.zeros(3);
        hp.le(hd.b.size(), 8);  // data segment size
        hp.le(1, 8);            // head of the free list: H5HL_FREE_NULL (no free block)
        align8();
        hp.le(pos_ + 32, 8);    // data segment address: right behind this 32-byte header
        hp.bytes(hd.b.data(), hd.b.size());
        g.heap = put_block(hp);
        // symbol table node, allocated at its full size 8 + 2K * 40
        Buf sn;
        sn.bytes("SNOD", 4);
        sn.u8(1);
        sn.u8(0);
        sn.le(names.size(), 2);
        uint64_t last_off = 0;
        for (auto &kv : names) {
            if (kv.second) {
                const Group &c = g.groups[kv.first];
                symbol_entry(sn, off[kv.first], c.ohdr, true, c.btree, c.heap);
            } else {
                symbol_entry(sn, off[kv.first], ds_addr[kv.first], false, 0, 0);
            }
            last_off = off[kv.first];
        }
        sn.zeros(8 + (size_t)2 * leaf_k_ * 40 - sn.b.size());
        const uint64_t snod = put_block(sn);
        // version-1 B-tree, group node, level 0, one child; allocated at its full size 24 + 2K * 8 + (2K + 1) * 8
        Buf bt;
        bt.bytes("TREE", 4);
        bt.u8(0);
        bt.u8(0);
        bt.le(names.empty() ? 0 : 1, 2);
        bt.le(UNDEF, 8);
        bt.le(UNDEF, 8);
        bt.le(0, 8);         // key 0: the empty string
        bt.le(snod, 8);      // child 0
        bt.le(last_off, 8);  // key 1: the largest name of child 0
        bt.zeros(24 + (size_t)2 * INTERNAL_K * 8 + (2 * INTERNAL_K + 1) * 8 - bt.b.size());
        g.btree = put_block(bt);
        // object header of the group: one Symbol Table message
        Buf h;
        h.u8(1);
        h.u8(0);
        h.le(1, 2);
        h.le(1, 4);
        h.le(24, 4);
        h.zeros(4);
        msg_header(h, 0x0011, 16, 0);
        h.le(g.btree, 8);
        h.le(g.heap, 8);
        g.ohdr = put_block(h);
    }
};

}  // namespace h5w
}  // namespace fans
