"""A small HDF5 writer and reader, enough for LOKI's on-disk formats.

libhdf5 and h5py are absent from this image, so the files are laid out by hand after the published HDF5 File Format
Specification (version 1.1 structures, what libhdf5 1.8 writes by default and every later libhdf5 reads): version-0
superblock, version-1 object headers, "old style" groups (a symbol-table message pointing at a version-1 B-tree of
symbol-table nodes plus a local heap of names), version-1 dataspace / datatype / attribute messages, version-3 data
layout (contiguous).  Those are the structures the reference's calls produce: H5Fcreate / H5Gcreate / H5Dcreate /
H5Acreate with default property lists (ReaderWriterBase.C:24-330, RestartWriter.C:55-470).

The reader walks the same structures and is pinned (tests/test_cpu_outputs.py) against files libhdf5 itself wrote:
scipy's MATLAB v7.3 fixture (tests/golden/libhdf5_written_*.mat: user block, symbol-table group, version-1/2 layout) and,
where the reference tree is present, the reference's own test/External2D/rho_init_*.h5 (compact new-style groups: Link
messages in a version-1 object header).  It also understands chunked layouts with the deflate filter and object-header
continuation blocks because such files are what a LOKI baseline would be.  PARITY OF THE WRITER IS UNPINNED against
libhdf5 itself (none in the image): its message bytes are compared with the genuine file's where they describe the
same thing, and everything it writes is read back by the pinned reader.

Datasets hold numpy arrays; the three element types LOKI writes are '<f8' (H5T_NATIVE_DOUBLE), '>i4' (H5T_STD_I32BE)
and 'u1' (H5T_NATIVE_UCHAR).  A dataset of shape () is a scalar dataspace (H5Screate(H5S_SCALAR))."""
import struct
import zlib

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF
LEAF_K = 4            # symbol-table node holds 2 * LEAF_K entries (H5Pset_sym_k default)
INTERNAL_K = 16       # B-tree node holds 2 * INTERNAL_K children (default)


class Dataset:
    def __init__(self, data, attrs=None, scalar=False):
        a = np.asarray(data)
        if a.dtype.kind == "f":
            a = a.astype("<f8", copy=False)
        elif a.dtype.kind == "i":
            a = a.astype(">i4")
        elif a.dtype.kind in "uSb":
            a = a.astype("u1", copy=False) if a.dtype.kind != "S" else np.frombuffer(a.tobytes(), dtype="u1")
        else:
            raise TypeError("unsupported element type %r" % a.dtype)
        if scalar:
            a = a.reshape(())
        self.data = a
        self.attrs = dict(attrs or {})

    @property
    def shape(self):
        return self.data.shape


class Group:
    def __init__(self):
        self.children = {}
        self.attrs = {}

    def group(self, name):
        g = self.children.get(name)
        if g is None:
            g = self.children[name] = Group()
        if not isinstance(g, Group):
            raise KeyError("%r is a dataset" % name)
        return g

    def put(self, name, data, scalar=False):
        if name in self.children:
            raise KeyError("%r exists already (H5Dcreate fails on an existing name)" % name)
        d = self.children[name] = data if isinstance(data, Dataset) else Dataset(data, scalar=scalar)
        return d

    def __getitem__(self, path):
        node = self
        for part in [p for p in path.split("/") if p]:
            node = node.children[part]
        return node

    def __contains__(self, name):
        return name in self.children

    def names(self):
        return list(self.children)


# ------------------------------------------------------------------ writer

def _pad8(b):
    return b + b"\0" * (-len(b) % 8)


def _dtype_message(dt):
    dt = np.dtype(dt)
    if dt.kind == "f" and dt.itemsize == 8:
        order = 0 if dt.byteorder in "<=|" else 1
        # class 1 (floating point) version 1; implied mantissa normalisation, sign bit 63
        return struct.pack("<BBBBI", 0x11, 0x20 | order, 63, 0, 8) + struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
    if dt.kind in "iu":
        order = 1 if dt.byteorder == ">" else 0
        signed = 0x08 if dt.kind == "i" else 0
        return struct.pack("<BBBBI", 0x10, order | signed, 0, 0, dt.itemsize) + struct.pack("<HH", 0, 8 * dt.itemsize)
    raise TypeError("unsupported element type %r" % dt)


def _dspace_message(shape):
    # version 1; rank 0 is the scalar dataspace
    return struct.pack("<BBBBI", 1, len(shape), 0, 0, 0) + b"".join(struct.pack("<Q", n) for n in shape)


def _attr_message(name, value):
    a = value.data if isinstance(value, Dataset) else Dataset(value).data
    nm = name.encode() + b"\0"
    dt, ds = _dtype_message(a.dtype), _dspace_message(a.shape)
    return (struct.pack("<BBHHH", 1, 0, len(nm), len(dt), len(ds)) + _pad8(nm) + _pad8(dt) + _pad8(ds) +
            np.ascontiguousarray(a).tobytes())


def _object_header(messages):
    body = b""
    for mtype, data in messages:
        data = _pad8(data)
        body += struct.pack("<HHBBBB", mtype, len(data), 0, 0, 0, 0) + data
    # version, reserved, message count, reference count, size of the message block; the block starts 8-aligned
    return struct.pack("<BBHII", 1, 0, len(messages), 1, len(body)) + b"\0" * 4 + body


class _Writer:
    def __init__(self, fh):
        self.fh = fh
        self.pos = 0
        self.internal_k = INTERNAL_K

    def alloc(self, size):
        self.pos += -self.pos % 8
        at = self.pos
        self.pos += size
        return at

    def put(self, at, data):
        self.fh.seek(at)
        self.fh.write(data)

    def plan(self, root):
        """largest group decides the B-tree rank so that one level-0 node addresses all its symbol-table nodes"""
        most = 0
        stack = [root]
        while stack:
            g = stack.pop()
            most = max(most, len(g.children))
            stack += [c for c in g.children.values() if isinstance(c, Group)]
        nodes = -(-most // (2 * LEAF_K))
        self.internal_k = max(INTERNAL_K, -(-nodes // 2))
        if self.internal_k > 0xFFFF:
            raise ValueError("group too large for one B-tree node")

    def write_dataset(self, d):
        a = d.data
        nbytes = a.size * a.itemsize
        addr = self.alloc(nbytes) if nbytes else UNDEF
        if nbytes:
            self.fh.seek(addr)
            c = a if a.flags.c_contiguous else np.ascontiguousarray(a)
            self.fh.write(memoryview(c).cast("B") if c.ndim else c.tobytes())
        msgs = [(0x0001, _dspace_message(a.shape)), (0x0003, _dtype_message(a.dtype)),
                (0x0005, struct.pack("<BBBB", 2, 2, 2, 0)),            # fill value v2: late allocation, if-set, undefined
                (0x0008, struct.pack("<BBQQ", 3, 1, addr, nbytes))]    # layout v3, contiguous
        msgs += [(0x000C, _attr_message(k, v)) for k, v in d.attrs.items()]
        oh = _object_header(msgs)
        at = self.alloc(len(oh))
        self.put(at, oh)
        return at

    def write_group(self, g):
        """returns (object header address, B-tree address, heap address)"""
        entries = []
        for name in sorted(g.children, key=lambda s: s.encode()):   # strcmp order, as H5G_node_cmp3 expects
            c = g.children[name]
            if isinstance(c, Group):
                entries.append((name, 1) + self.write_group(c))
            else:
                entries.append((name, 0, self.write_dataset(c), 0, 0))
        # local heap: the empty string at offset 0, then the names, each padded to 8 bytes
        heap = bytearray(8)
        offs = []
        for e in entries:
            offs.append(len(heap))
            heap += _pad8(e[0].encode() + b"\0")
        free_at = len(heap)
        seg_size = max(88, free_at + 16)        # libhdf5's initial heap size is 88 bytes for a new group
        seg_size += -seg_size % 8
        heap += struct.pack("<QQ", 1, seg_size - free_at) + b"\0" * (seg_size - free_at - 16)
        heap_at = self.alloc(32)
        seg_at = self.alloc(seg_size)
        self.put(heap_at, b"HEAP" + struct.pack("<BBBBQQQ", 0, 0, 0, 0, seg_size, free_at, seg_at))
        self.put(seg_at, bytes(heap))
        # symbol-table nodes
        per = 2 * LEAF_K
        snods, keys = [], [0]
        for i in range(0, len(entries), per):
            chunk = entries[i:i + per]
            body = b"SNOD" + struct.pack("<BBH", 1, 0, len(chunk))
            for j, (name, cache, oh, bt, hp) in enumerate(chunk):
                scratch = struct.pack("<QQ", bt, hp) if cache == 1 else b"\0" * 16
                body += struct.pack("<QQII", offs[i + j], oh, cache, 0) + scratch
            body += b"\0" * (8 + per * 40 - len(body))
            at = self.alloc(len(body))
            self.put(at, body)
            snods.append(at)
            keys.append(offs[i + len(chunk) - 1])
        k2 = 2 * self.internal_k
        assert len(snods) <= k2
        tree = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(snods), UNDEF, UNDEF)
        for i, at in enumerate(snods):
            tree += struct.pack("<QQ", keys[i], at)
        tree += struct.pack("<Q", keys[len(snods)])
        tree += b"\0" * (24 + (k2 + 1) * 8 + k2 * 8 - len(tree))
        tree_at = self.alloc(len(tree))
        self.put(tree_at, tree)
        msgs = [(0x0011, struct.pack("<QQ", tree_at, heap_at))]
        msgs += [(0x000C, _attr_message(k, v)) for k, v in g.attrs.items()]
        oh = _object_header(msgs)
        oh_at = self.alloc(len(oh))
        self.put(oh_at, oh)
        return oh_at, tree_at, heap_at


def write(path, root):
    """lay `root` (a Group) out as an HDF5 file; raw data is streamed from the arrays, nothing is copied twice"""
    with open(path, "wb") as fh:
        w = _Writer(fh)
        w.plan(root)
        w.alloc(96)                                   # superblock, written last (it holds the end-of-file address)
        oh_at, tree_at, heap_at = w.write_group(root)
        eof = w.pos + (-w.pos % 8)
        sb = SIGNATURE + struct.pack("<BBBBBBBB", 0, 0, 0, 0, 0, 8, 8, 0)
        sb += struct.pack("<HHI", LEAF_K, w.internal_k, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
        sb += struct.pack("<QQII", 0, oh_at, 1, 0) + struct.pack("<QQ", tree_at, heap_at)
        assert len(sb) == 96
        fh.seek(0)
        fh.write(sb)
        fh.truncate(eof)


# ------------------------------------------------------------------ reader

class FormatError(ValueError):
    pass


class _Reader:
    def __init__(self, buf):
        self.b = buf
        at = 0
        while buf[at:at + 8] != SIGNATURE:             # the superblock sits at 0, 512, 1024, ... (user block)
            at = 512 if at == 0 else 2 * at
            if at + 8 > len(buf):
                raise FormatError("not an HDF5 file")
        ver = buf[at + 8]
        if ver not in (0, 1):
            raise FormatError("superblock version %d (only the 0 / 1 layout is read)" % ver)
        so, sl = buf[at + 13], buf[at + 14]
        if (so, sl) != (8, 8):
            raise FormatError("offsets / lengths of %d / %d bytes" % (so, sl))
        self.leaf_k, self.internal_k = struct.unpack_from("<HH", buf, at + 16)
        p = at + 24 + (4 if ver == 1 else 0)
        self.base, _, self.eof, _ = struct.unpack_from("<QQQQ", buf, p)
        if self.base == 0 and at:
            self.base = 0                               # addresses relative to the file start, as written
        self.root_entry = self.sym_entry(p + 32)

    def sym_entry(self, p):
        name_off, oh, cache, _ = struct.unpack_from("<QQII", self.b, p)
        bt, hp = struct.unpack_from("<QQ", self.b, p + 24)
        return dict(name_off=name_off, oh=oh, cache=cache, btree=bt, heap=hp)

    def messages(self, addr):
        """all messages of a version-1 object header, continuation blocks followed"""
        b = self.b
        p = self.base + addr
        ver, _, nmsg, _, size = struct.unpack_from("<BBHII", b, p)
        if ver != 1:
            raise FormatError("object header version %d at %#x" % (ver, addr))
        blocks = [(p + 16, size)]
        out = []
        while blocks:
            q, left = blocks.pop(0)
            while left >= 8 and len(out) < nmsg:
                mtype, msize, flags = struct.unpack_from("<HHB", b, q)
                data = b[q + 8:q + 8 + msize]
                if mtype == 0x0010:
                    ca, cl = struct.unpack_from("<QQ", data)
                    blocks.append((self.base + ca, cl))
                out.append((mtype, flags, data))
                q += 8 + msize
                left -= 8 + msize
        return out

    def heap_name(self, heap_addr, off):
        p = self.base + heap_addr
        if self.b[p:p + 4] != b"HEAP":
            raise FormatError("local heap signature")
        seg = struct.unpack_from("<Q", self.b, p + 24)[0]
        s = self.base + seg + off
        raw = bytes(self.b[s:s + 1024])
        return raw[:raw.index(b"\0")].decode()

    def group_entries(self, btree, heap):
        out = []

        def node(addr):
            p = self.base + addr
            if self.b[p:p + 4] == b"SNOD":
                n = struct.unpack_from("<H", self.b, p + 6)[0]
                for i in range(n):
                    e = self.sym_entry(p + 8 + 40 * i)
                    out.append((self.heap_name(heap, e["name_off"]), e))
                return
            if self.b[p:p + 4] != b"TREE":
                raise FormatError("B-tree node signature at %#x" % addr)
            ntype, level, used = struct.unpack_from("<BBH", self.b, p + 4)
            if ntype != 0:
                raise FormatError("group B-tree expected")
            q = p + 24
            for i in range(used):
                child = struct.unpack_from("<Q", self.b, q + 8)[0]
                node(child)
                q += 16
        node(btree)
        return out

    @staticmethod
    def parse_dtype(data):
        cv, b0, b1, b2, size = struct.unpack_from("<BBBBI", data)
        cls = cv & 0x0F
        if cls == 0:
            order = ">" if b0 & 1 else "<"
            kind = "i" if b0 & 0x08 else "u"
            return np.dtype("%s%s%d" % (order if size > 1 else "|", kind, size)) if size > 1 else np.dtype(kind + "1")
        if cls == 1:
            return np.dtype(("%sf%d") % (">" if b0 & 1 else "<", size))
        if cls == 3:
            return np.dtype("S%d" % size)
        if cls == 7:
            return np.dtype("<u8") if size == 8 else np.dtype("V%d" % size)
        return np.dtype("V%d" % size)

    @staticmethod
    def parse_dspace(data):
        ver, rank, flags = struct.unpack_from("<BBB", data)
        p = 8 if ver == 1 else 4
        if ver == 2 and data[3] == 2:
            return None                                  # null dataspace
        return tuple(struct.unpack_from("<%dQ" % rank, data, p)) if rank else ()

    def chunked(self, btree_addr, shape, chunk, dt, filters):
        out = np.zeros(shape, dtype=dt)
        rank = len(shape)

        def node(addr):
            p = self.base + addr
            if self.b[p:p + 4] != b"TREE":
                raise FormatError("chunk B-tree signature")
            ntype, level, used = struct.unpack_from("<BBH", self.b, p + 4)
            q = p + 24
            ksz = 8 + 8 * (rank + 1)
            for i in range(used):
                csize, mask = struct.unpack_from("<II", self.b, q)
                offs = struct.unpack_from("<%dQ" % (rank + 1), self.b, q + 8)
                child = struct.unpack_from("<Q", self.b, q + ksz)[0]
                if level:
                    node(child)
                else:
                    raw = bytes(self.b[self.base + child:self.base + child + csize])
                    for k, fid in reversed(list(enumerate(filters))):
                        if mask & (1 << k):
                            continue
                        if fid == 1:
                            raw = zlib.decompress(raw)
                        elif fid == 2:
                            n = dt.itemsize
                            raw = np.frombuffer(raw, "u1").reshape(n, -1).T.tobytes()
                        else:
                            raise FormatError("filter %d" % fid)
                    blk = np.frombuffer(raw, dtype=dt).reshape(chunk)
                    sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, chunk, shape))
                    out[sl] = blk[tuple(slice(0, s.stop - s.start) for s in sl)]
                q += ksz + 8
        if btree_addr != UNDEF:
            node(btree_addr)
        return out

    def attr(self, data):
        ver = data[0]
        if ver == 1:
            _, _, nsz, tsz, ssz = struct.unpack_from("<BBHHH", data)
            p = 8
            name = bytes(data[p:p + nsz]).split(b"\0")[0].decode()
            p += nsz + (-nsz % 8)
            dt = self.parse_dtype(data[p:p + tsz])
            p += tsz + (-tsz % 8)
            shape = self.parse_dspace(data[p:p + ssz])
            p += ssz + (-ssz % 8)
        else:
            raise FormatError("attribute message version %d" % ver)
        n = int(np.prod(shape)) if shape else 1
        val = np.frombuffer(bytes(data[p:p + n * dt.itemsize]), dtype=dt).reshape(shape or ())
        return name, val

    def load(self, entry):
        msgs = self.messages(entry["oh"])
        types = {m[0] for m in msgs}
        attrs = dict(self.attr(d) for t, _, d in msgs if t == 0x000C)
        if 0x0011 in types or entry["cache"] == 1:
            g = Group()
            g.attrs = attrs
            bt, hp = entry["btree"], entry["heap"]
            for t, _, d in msgs:
                if t == 0x0011:
                    bt, hp = struct.unpack_from("<QQ", d)
            for name, e in self.group_entries(bt, hp):
                g.children[name] = self.load(e)
            return g
        if 0x0006 in types or 0x0002 in types:
            # a "new style" group with compact link storage: one Link message per member in the object header (what
            # h5py / libhdf5 >= 1.8 write for small groups unless told to stay with symbol tables)
            g = Group()
            g.attrs = attrs
            for t, _, d in msgs:
                if t == 0x0002 and len(d) >= 18:
                    flags = d[1]
                    p = 2 + (8 if flags & 1 else 0)
                    if struct.unpack_from("<Q", d, p)[0] != UNDEF:
                        raise FormatError("dense link storage (fractal heap) is not read")
                if t != 0x0006:
                    continue
                flags = d[1]
                p = 2
                ltype = 0
                if flags & 0x08:
                    ltype = d[p]
                    p += 1
                if flags & 0x04:
                    p += 8
                if flags & 0x10:
                    p += 1
                nlen_size = 1 << (flags & 3)
                nlen = int.from_bytes(bytes(d[p:p + nlen_size]), "little")
                p += nlen_size
                name = bytes(d[p:p + nlen]).decode()
                p += nlen
                if ltype != 0:
                    continue                                   # soft / external links carry no data
                addr = struct.unpack_from("<Q", d, p)[0]
                g.children[name] = self.load(dict(name_off=0, oh=addr, cache=0, btree=0, heap=0))
            return g
        dt = shape = None
        layout = None
        filters = []
        for t, _, d in msgs:
            if t == 0x0001:
                shape = self.parse_dspace(d)
            elif t == 0x0003:
                dt = self.parse_dtype(d)
            elif t == 0x0008:
                layout = d
            elif t == 0x000B:
                nf = d[1]
                p = 8 if d[0] == 1 else 2
                for _ in range(nf):
                    fid, nlen, fl, ncd = struct.unpack_from("<HHHH", d, p)
                    p += 8
                    if d[0] == 1 or fid >= 256:
                        p += nlen + (-nlen % 8 if d[0] == 1 else 0)
                    p += 4 * ncd
                    if d[0] == 1 and ncd % 2:
                        p += 4
                    filters.append(fid)
        if dt is None or shape is None or layout is None:
            raise FormatError("object at %#x is neither a group nor a dataset" % entry["oh"])
        n = int(np.prod(shape)) if shape else 1
        if layout[0] in (1, 2):
            # versions 1 / 2: rank, class, 5 reserved bytes, address (absent for compact), 4-byte dimensions
            rank1, cls = layout[1], layout[2]
            p = 8
            addr = UNDEF
            if cls != 0:
                addr = struct.unpack_from("<Q", layout, p)[0]
                p += 8
            dims = struct.unpack_from("<%dI" % rank1, layout, p)
            p += 4 * rank1
            if cls == 2:
                arr = self.chunked(addr, shape, dims[:-1] if len(dims) > len(shape) else dims, dt, filters)
            elif cls == 1:
                arr = (np.zeros(shape, dtype=dt) if addr == UNDEF else
                       np.frombuffer(self.b, dtype=dt, count=n, offset=self.base + addr).reshape(shape))
            else:
                size = struct.unpack_from("<I", layout, p)[0]
                arr = np.frombuffer(bytes(layout[p + 4:p + 4 + size]), dtype=dt, count=n).reshape(shape)
            ds = Dataset.__new__(Dataset)
            ds.data, ds.attrs = arr, attrs
            return ds
        if layout[0] != 3:
            raise FormatError("data layout message version %d" % layout[0])
        cls = layout[1]
        if cls == 1:
            addr, size = struct.unpack_from("<QQ", layout, 2)
            if addr == UNDEF:
                arr = np.zeros(shape, dtype=dt)
            else:
                arr = np.frombuffer(self.b, dtype=dt, count=n, offset=self.base + addr).reshape(shape)
        elif cls == 0:
            size = struct.unpack_from("<H", layout, 2)[0]
            arr = np.frombuffer(bytes(layout[4:4 + size]), dtype=dt, count=n).reshape(shape)
        elif cls == 2:
            rank1 = layout[2]
            bt = struct.unpack_from("<Q", layout, 3)[0]
            dims = struct.unpack_from("<%dI" % rank1, layout, 11)
            arr = self.chunked(bt, shape, dims[:-1], dt, filters)
        else:
            raise FormatError("layout class %d" % cls)
        ds = Dataset.__new__(Dataset)
        ds.data, ds.attrs = arr, attrs
        return ds


def read(path):
    """the file's root Group; contiguous datasets are views of one read-only memory map of the file"""
    buf = np.memmap(path, dtype="u1", mode="r")
    r = _Reader(memoryview(buf))
    return r.load(r.root_entry)
