"""Independent pure-Python reader for the subset of HDF5 that fans_b200/host/h5write.hpp emits, written from the HDF5 File Format
Specification (superblock v0, symbol-table groups, v1 object headers, contiguous datasets, v1 attributes).  It is the checker of
the writer in the CPU tests; it also walks the new-style compact groups of the reference's own sphere32.h5, so the same code is
exercised against a file produced by the real HDF5 library (chunked datasets are outside its scope)."""
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class H5File:
    def __init__(self, path):
        self.b = open(path, "rb").read()
        b = self.b
        assert b[:8] == b"\x89HDF\r\n\x1a\n", "not an HDF5 file"
        assert b[8] == 0, "superblock version %d" % b[8]
        assert b[13] == 8 and b[14] == 8, "only 8-byte offsets / lengths"
        self.leaf_k, self.internal_k = struct.unpack("<HH", b[16:20])
        base, _, self.eof, _ = struct.unpack("<QQQQ", b[24:56])
        assert base == 0 and self.eof == len(b), "end-of-file address %d != file size %d" % (self.eof, len(b))
        _, self.root, cache, _ = struct.unpack("<QQII", b[56:80])
        self.root_scratch = struct.unpack("<QQ", b[80:96]) if cache == 1 else None

    def u(self, p, n):
        return int.from_bytes(self.b[p:p + n], "little")

    def messages(self, addr):
        b = self.b
        assert b[addr] == 1, "object header version %d" % b[addr]
        assert addr % 8 == 0
        n, size = self.u(addr + 2, 2), self.u(addr + 8, 4)
        blocks, out, seen, bi = [(addr + 16, size)], [], 0, 0
        while bi < len(blocks) and seen < n:
            p, sz = blocks[bi]
            end = p + sz
            while p + 8 <= end and seen < n:
                t, s = self.u(p, 2), self.u(p + 2, 2)
                assert s % 8 == 0, "message data not padded to 8 bytes"
                out.append((t, p + 8, s))
                if t == 0x10:
                    blocks.append((self.u(p + 8, 8), self.u(p + 16, 8)))
                p += 8 + s
                seen += 1
            bi += 1
        assert seen == n, "object header announces %d messages, found %d" % (n, seen)
        return out

    def children(self, addr):
        """{name: (object header address, is_group)} of the group at addr"""
        out = {}
        for t, p, s in self.messages(addr):
            if t == 0x11:   # symbol table message
                btree, heap = self.u(p, 8), self.u(p + 8, 8)
                assert self.b[heap:heap + 4] == b"HEAP" and self.b[heap + 4] == 0
                dsize, free, daddr = self.u(heap + 8, 8), self.u(heap + 16, 8), self.u(heap + 24, 8)
                assert free == 1 or free < dsize, "bad free-list head"
                assert self.b[daddr] == 0, "heap offset 0 must hold the empty string"
                names_seen = []
                self._btree(btree, daddr, dsize, out, names_seen)
                assert names_seen == sorted(names_seen), "symbol table entries are not in name order"
            elif t == 0x06:  # link message (compact new-style group)
                q = p
                assert self.b[q] == 1
                flags = self.b[q + 1]
                q += 2
                ltype = 0
                if flags & 0x08:
                    ltype = self.b[q]
                    q += 1
                if flags & 0x04:
                    q += 8
                if flags & 0x10:
                    q += 1
                lsz = 1 << (flags & 3)
                nlen = self.u(q, lsz)
                q += lsz
                name = self.b[q:q + nlen].decode()
                q += nlen
                if ltype == 0:
                    out[name] = (self.u(q, 8), None)
        return out

    def _btree(self, node, heap_data, heap_size, out, names_seen):
        b = self.b
        if b[node:node + 4] == b"TREE":
            assert b[node + 4] == 0, "not a group B-tree"
            used = self.u(node + 6, 2)
            assert used <= 2 * self.internal_k
            assert node + 24 + 2 * self.internal_k * 8 + (2 * self.internal_k + 1) * 8 <= len(b), "B-tree node is not allocated at full size"
            p = node + 24
            for _ in range(used):
                p += 8   # key
                self._btree(self.u(p, 8), heap_data, heap_size, out, names_seen)
                p += 8
        else:
            assert b[node:node + 4] == b"SNOD" and b[node + 4] == 1
            assert node + 8 + 2 * self.leaf_k * 40 <= len(b), "symbol table node is not allocated at full size"
            n = self.u(node + 6, 2)
            assert n <= 2 * self.leaf_k
            for i in range(n):
                p = node + 8 + 40 * i
                noff, ohdr, cache = self.u(p, 8), self.u(p + 8, 8), self.u(p + 16, 4)
                assert noff < heap_size
                end = b.index(b"\0", heap_data + noff)
                name = b[heap_data + noff:end].decode()
                names_seen.append(name.encode())
                out[name] = (ohdr, cache == 1)

    def dataset(self, addr):
        """(array or None if not a contiguous dataset, attrs)"""
        dims = dtype = layout = None
        attrs = {}
        for t, p, s in self.messages(addr):
            if t == 0x01:
                ver, rank, flags = self.b[p], self.b[p + 1], self.b[p + 2]
                assert ver == 1
                dims = [self.u(p + 8 + 8 * i, 8) for i in range(rank)]
            elif t == 0x03:
                dtype = self._dtype(p)
            elif t == 0x08:
                assert self.b[p] == 3, "layout version"
                layout = (self.b[p + 1], self.u(p + 2, 8), self.u(p + 10, 8))
            elif t == 0x0C:
                assert self.b[p] == 1
                nlen, dtl, dsl = self.u(p + 2, 2), self.u(p + 4, 2), self.u(p + 6, 2)
                q = p + 8
                name = self.b[q:q + nlen].rstrip(b"\0").decode()
                q += (nlen + 7) // 8 * 8
                adt = self._dtype(q)
                q += (dtl + 7) // 8 * 8
                assert self.b[q] == 1 and self.b[q + 1] == 0, "only scalar attributes"
                q += (dsl + 7) // 8 * 8
                if adt[0] == "S":
                    attrs[name] = self.b[q:q + adt[1]].split(b"\0")[0].decode()
                else:   # scalar integer attribute (num_crystals / num_GB of a grain-boundary image)
                    attrs[name] = int(np.frombuffer(self.b, dtype=adt, count=1, offset=q)[0])
        if dims is None or dtype is None or layout is None:
            return None, attrs
        if layout[0] != 1:
            return None, attrs
        n = int(np.prod(dims)) if dims else 1
        assert layout[2] == n * np.dtype(dtype).itemsize, "layout size does not match dataspace x datatype"
        assert layout[1] + layout[2] <= len(self.b)
        return np.frombuffer(self.b, dtype=dtype, count=n, offset=layout[1]).reshape(dims), attrs

    def _dtype(self, p):
        cls, ver = self.b[p] & 15, self.b[p] >> 4
        assert ver == 1
        size = self.u(p + 4, 4)
        bits0 = self.b[p + 1]
        if cls == 0:
            assert bits0 & 1 == 0, "big endian"
            assert self.u(p + 8, 2) == 0 and self.u(p + 10, 2) == 8 * size
            return ("<i%d" if bits0 & 8 else "<u%d") % size
        if cls == 1:
            assert bits0 & 1 == 0 and (bits0 >> 4) & 3 == 2
            prec, eloc, esz, mloc, msz, bias = self.u(p + 10, 2), self.b[p + 12], self.b[p + 13], self.b[p + 14], self.b[p + 15], self.u(p + 16, 4)
            assert (size, prec, self.b[p + 2], eloc, esz, mloc, msz, bias) in ((8, 64, 63, 52, 11, 0, 52, 1023), (4, 32, 31, 23, 8, 0, 23, 127)), "not IEEE"
            return "<f%d" % size
        if cls == 3:
            return ("S", size)
        raise AssertionError("datatype class %d" % cls)

    def walk(self, addr=None, prefix=""):
        """{path: (array, attrs)} of every dataset below addr"""
        addr = self.root if addr is None else addr
        out = {}
        for name, (ohdr, is_group) in self.children(addr).items():
            types = [t for t, _, _ in self.messages(ohdr)]
            if 0x11 in types or (0x02 in types and 0x01 not in types):
                out.update(self.walk(ohdr, prefix + "/" + name))
            else:
                out[prefix + "/" + name] = self.dataset(ohdr)
        return out
