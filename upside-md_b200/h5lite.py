"""Minimal self-contained HDF5 reader / writer (numpy only, no libhdf5, no h5py).

Why it exists: the image has neither libhdf5 nor h5py/PyTables, yet the drop-in boundary of this
project is the HDF5 ``.up`` configuration written by the reference's ``py/upside_config.py`` and the
HDF5 parameter libraries under ``parameters/``.  This module reads what libhdf5 1.8 / PyTables emit
for such files (superblock v0, object header v1, symbol-table groups, contiguous / compact / chunked
datasets with deflate + shuffle + fletcher32, v1-v3 attribute messages) and writes the same subset.

The C++ engine has its own reader (csrc/h5lite.cpp); the two are cross-checked in tests/test_h5lite.py.

Model: ``File`` -> nested ``Group`` (dict of children + attrs) -> ``Dataset`` (numpy array + attrs).
"""
import struct
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIGNATURE = b"\x89HDF\r\n\x1a\n"


class Dataset:
    def __init__(self, data, attrs=None, chunks=None, compress=False):
        self.data = np.asarray(data)
        self.attrs = dict(attrs or {})
        self.chunks = chunks        # writer hint: tuple -> chunked layout
        self.compress = compress    # writer hint: shuffle + deflate + fletcher32 (needs chunks)

    @property
    def shape(self):
        return self.data.shape

    def __getitem__(self, k):
        return self.data[k]


class Group:
    def __init__(self):
        self.children = {}
        self.attrs = {}

    # dict-like helpers -------------------------------------------------------------------------
    def __contains__(self, path):
        try:
            self[path]
            return True
        except KeyError:
            return False

    def __getitem__(self, path):
        node = self
        for part in [p for p in path.split("/") if p and p != "."]:
            if not isinstance(node, Group):
                raise KeyError(path)
            node = node.children[part]
        return node

    def keys(self):
        return sorted(self.children)

    def create_group(self, path):
        node = self
        for part in [p for p in path.split("/") if p]:
            if part not in node.children:
                node.children[part] = Group()
            node = node.children[part]
        return node

    def create_dataset(self, path, data, attrs=None, chunks=None, compress=False):
        parts = [p for p in path.split("/") if p]
        g = self.create_group("/".join(parts[:-1])) if len(parts) > 1 else self
        d = Dataset(data, attrs, chunks, compress)
        g.children[parts[-1]] = d
        return d


class File(Group):
    pass


# ================================================================================================
# reader
# ================================================================================================

class _Reader:
    def __init__(self, buf):
        self.b = buf
        if buf[:8] != SIGNATURE:
            raise ValueError("not an HDF5 file")
        ver = buf[8]
        if ver not in (0, 1):
            raise ValueError("only superblock v0/v1 supported (got %d)" % ver)
        if buf[13] != 8 or buf[14] != 8:
            raise ValueError("only 8-byte offsets/lengths supported")
        off = 24 + (4 if ver == 1 else 0)
        self.base, _fs, self.eof, _drv = struct.unpack_from("<QQQQ", buf, off)
        off += 32
        _name_off, self.root_ohdr, _cache, _r = struct.unpack_from("<QQII", buf, off)

    # -- object headers -------------------------------------------------------------------------
    def messages(self, addr):
        b = self.b
        ver, _r, nmsg, _refc, hsize = struct.unpack_from("<BBHII", b, addr)
        if ver != 1:
            raise ValueError("only v1 object headers supported (got %d at 0x%x)" % (ver, addr))
        blocks = [(addr + 16, hsize)]
        out = []
        bi = 0
        while bi < len(blocks) and len(out) < nmsg:
            p, ln = blocks[bi]
            end = p + ln
            while p + 8 <= end and len(out) < nmsg:
                mtype, msize, mflags = struct.unpack_from("<HHB", b, p)
                body = b[p + 8:p + 8 + msize]
                p += 8 + msize
                if mtype == 0x10:
                    coff, clen = struct.unpack_from("<QQ", body, 0)
                    blocks.append((coff, clen))
                out.append((mtype, mflags, body))
            bi += 1
        return out

    @staticmethod
    def parse_dtype(body):
        cv, bf0, bf1, bf2, size = struct.unpack_from("<BBBBI", body, 0)
        cls = cv & 0xF
        if cls == 0:
            if bf0 & 1:
                raise ValueError("big-endian integers unsupported")
            return np.dtype("<%s%d" % ("i" if bf0 & 8 else "u", size)), 8 + 4
        if cls == 1:
            if bf0 & 1:
                raise ValueError("big-endian floats unsupported")
            return np.dtype("<f%d" % size), 8 + 12
        if cls == 3:
            return np.dtype("S%d" % size), 8
        raise ValueError("unsupported datatype class %d" % cls)

    @staticmethod
    def parse_space(body):
        ver = body[0]
        rank = body[1]
        flags = body[2]
        if ver == 1:
            off = 8
        elif ver == 2:
            off = 4
            if body[3] == 2:   # null dataspace
                return None
        else:
            raise ValueError("dataspace version %d" % ver)
        dims = struct.unpack_from("<%dQ" % rank, body, off) if rank else ()
        return tuple(int(d) for d in dims)

    def parse_attr(self, body):
        ver = body[0]
        name_sz, dt_sz, sp_sz = struct.unpack_from("<HHH", body, 2)
        p = 8
        if ver == 3:
            p = 9
        pad = (lambda n: (n + 7) & ~7) if ver == 1 else (lambda n: n)
        name = body[p:p + name_sz].split(b"\0")[0].decode()
        p += pad(name_sz)
        dt, _ = self.parse_dtype(body[p:p + dt_sz])
        p += pad(dt_sz)
        shape = self.parse_space(body[p:p + sp_sz])
        p += pad(sp_sz)
        if shape is None:
            return name, None
        n = int(np.prod(shape, dtype=np.int64)) if shape else 1
        arr = np.frombuffer(body, dtype=dt, count=n, offset=p).reshape(shape).copy()
        return name, (arr if shape else arr[()])

    # -- datasets -------------------------------------------------------------------------------
    def read_chunk_btree(self, addr, rank1, out):
        b = self.b
        if b[addr:addr + 4] != b"TREE":
            raise ValueError("bad chunk B-tree node")
        ntype, level, nent = struct.unpack_from("<BBH", b, addr + 4)
        p = addr + 24
        keysz = 8 + 8 * rank1
        for _ in range(nent):
            csize, fmask = struct.unpack_from("<II", b, p)
            offs = struct.unpack_from("<%dQ" % rank1, b, p + 8)
            child = struct.unpack_from("<Q", b, p + keysz)[0]
            p += keysz + 8
            if level > 0:
                self.read_chunk_btree(child, rank1, out)
            else:
                out.append((offs[:-1], csize, fmask, child))

    def read_dataset(self, msgs):
        dt = shape = None
        layout = None
        filters = []
        attrs = {}
        for mtype, _f, body in msgs:
            if mtype == 0x1:
                shape = self.parse_space(body)
            elif mtype == 0x3:
                dt, _ = self.parse_dtype(body)
            elif mtype == 0x8:
                layout = body
            elif mtype == 0xB:
                filters = self.parse_filters(body)
            elif mtype == 0xC:
                k, v = self.parse_attr(body)
                attrs[k] = v
        if layout[0] != 3:
            raise ValueError("only layout message v3 supported (got %d)" % layout[0])
        n = int(np.prod(shape, dtype=np.int64)) if shape else 1
        lclass = layout[1]
        chunks = None
        if lclass == 0:
            sz = struct.unpack_from("<H", layout, 2)[0]
            arr = np.frombuffer(layout, dtype=dt, count=n, offset=4).reshape(shape).copy()
        elif lclass == 1:
            addr, sz = struct.unpack_from("<QQ", layout, 2)
            if addr == UNDEF or n == 0:
                arr = np.zeros(shape, dtype=dt)
            else:
                arr = np.frombuffer(self.b, dtype=dt, count=n, offset=addr).reshape(shape).copy()
        elif lclass == 2:
            rank1 = layout[2]
            baddr = struct.unpack_from("<Q", layout, 3)[0]
            cdims = struct.unpack_from("<%dI" % rank1, layout, 11)
            chunks = tuple(cdims[:-1])
            arr = np.zeros(shape, dtype=dt)
            if baddr != UNDEF and n:
                recs = []
                self.read_chunk_btree(baddr, rank1, recs)
                for offs, csize, fmask, caddr in recs:
                    raw = bytes(self.b[caddr:caddr + csize])
                    for i, (fid, cd) in reversed(list(enumerate(filters))):
                        if fmask & (1 << i):
                            continue
                        if fid == 3:
                            raw = raw[:-4]
                        elif fid == 1:
                            raw = zlib.decompress(raw)
                        elif fid == 2:
                            es = dt.itemsize
                            ne = len(raw) // es
                            raw = np.frombuffer(raw, np.uint8, ne * es).reshape(es, ne).T.tobytes() + raw[ne * es:]
                        else:
                            raise ValueError("unsupported filter %d" % fid)
                    ch = np.frombuffer(raw, dtype=dt, count=int(np.prod(chunks))).reshape(chunks)
                    sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, chunks, shape))
                    sub = tuple(slice(0, s.stop - s.start) for s in sl)
                    arr[sl] = ch[sub]
        else:
            raise ValueError("layout class %d" % lclass)
        d = Dataset(arr, attrs, chunks, bool(filters))
        return d

    @staticmethod
    def parse_filters(body):
        ver, nf = body[0], body[1]
        p = 8 if ver == 1 else 2
        out = []
        for _ in range(nf):
            fid = struct.unpack_from("<H", body, p)[0]
            p += 2
            if ver == 1 or fid >= 256:
                nlen = struct.unpack_from("<H", body, p)[0]
                p += 2
            else:
                nlen = 0
            _flags, ncd = struct.unpack_from("<HH", body, p)
            p += 4
            p += ((nlen + 7) & ~7) if ver == 1 else nlen
            cd = struct.unpack_from("<%dI" % ncd, body, p)
            p += 4 * ncd
            if ver == 1 and ncd % 2:
                p += 4
            out.append((fid, cd))
        return out

    # -- groups ---------------------------------------------------------------------------------
    def group_entries(self, btree, heap):
        b = self.b
        if b[heap:heap + 4] != b"HEAP":
            raise ValueError("bad local heap")
        hdata = struct.unpack_from("<Q", b, heap + 24)[0]
        out = []

        def walk(addr):
            if b[addr:addr + 4] == b"TREE":
                _t, level, nent = struct.unpack_from("<BBH", b, addr + 4)
                p = addr + 24
                for _ in range(nent):
                    child = struct.unpack_from("<Q", b, p + 8)[0]
                    p += 16
                    walk(child)
            elif b[addr:addr + 4] == b"SNOD":
                nsym = struct.unpack_from("<H", b, addr + 6)[0]
                p = addr + 8
                for _ in range(nsym):
                    noff, ohdr = struct.unpack_from("<QQ", b, p)
                    s = hdata + noff
                    e = b.index(b"\0", s)
                    out.append((bytes(b[s:e]).decode(), ohdr))
                    p += 40
            else:
                raise ValueError("bad group node at 0x%x" % addr)
        walk(btree)
        return out

    def read_object(self, addr):
        msgs = self.messages(addr)
        types = [m[0] for m in msgs]
        if 0x11 in types:
            g = Group()
            for mtype, _f, body in msgs:
                if mtype == 0x11:
                    bt, hp = struct.unpack_from("<QQ", body, 0)
                    for name, ohdr in self.group_entries(bt, hp):
                        g.children[name] = self.read_object(ohdr)
                elif mtype == 0xC:
                    k, v = self.parse_attr(body)
                    g.attrs[k] = v
            return g
        if 0x8 in types:
            return self.read_dataset(msgs)
        if 0x2 in types or 0x6 in types:
            raise ValueError("new-style (link message) groups unsupported")
        g = Group()
        return g


def load(path):
    """Parse an HDF5 file fully into memory and return its root ``File`` group."""
    with open(path, "rb") as fh:
        buf = fh.read()
    r = _Reader(memoryview(buf).toreadonly() if False else buf)
    root = r.read_object(r.root_ohdr)
    f = File()
    f.children, f.attrs = root.children, root.attrs
    return f


# ================================================================================================
# writer
# ================================================================================================

def _pad8(b):
    return b + b"\0" * (-len(b) % 8)


def _dtype_msg(dt):
    dt = np.dtype(dt)
    if dt.kind in "iu":
        bf0 = 0x08 if dt.kind == "i" else 0
        return struct.pack("<BBBBI", 0x10 | 0, bf0, 0, 0, dt.itemsize) + struct.pack("<HH", 0, 8 * dt.itemsize)
    if dt.kind == "f":
        if dt.itemsize == 4:
            props = struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
            sign = 31
        elif dt.itemsize == 8:
            props = struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
            sign = 63
        else:
            raise ValueError("float size")
        return struct.pack("<BBBBI", 0x10 | 1, 0x20, sign, 0, dt.itemsize) + props
    if dt.kind == "S":
        return struct.pack("<BBBBI", 0x10 | 3, 0x00, 0, 0, max(dt.itemsize, 1))
    raise ValueError("unsupported dtype %r" % dt)


def _space_msg(shape, maxshape=None):
    if shape == ():
        return struct.pack("<BBBBI", 1, 0, 0, 0, 0)
    flags = 1 if maxshape is not None else 0
    out = struct.pack("<BBBBI", 1, len(shape), flags, 0, 0) + struct.pack("<%dQ" % len(shape), *shape)
    if maxshape is not None:
        out += struct.pack("<%dQ" % len(shape), *[UNDEF if m is None else m for m in maxshape])
    return out


def _normalise(value):
    a = np.asarray(value)
    if a.dtype.kind == "U":
        a = np.char.encode(a, "ascii")
    if a.dtype.kind == "O":
        raise ValueError("object arrays unsupported")
    if a.dtype.kind == "b":
        a = a.astype(np.int8)
    if a.dtype.byteorder == ">":
        a = a.astype(a.dtype.newbyteorder("<"))
    return a


def _attr_msg(name, value):
    a = _normalise(value)
    nm = name.encode() + b"\0"
    dt = _dtype_msg(a.dtype)
    sp = _space_msg(a.shape)
    body = struct.pack("<BBHHH", 1, 0, len(nm), len(dt), len(sp))
    body += _pad8(nm) + _pad8(dt) + _pad8(sp) + np.ascontiguousarray(a).tobytes()
    return 0xC, body


def _fletcher32(data):
    """HDF5's H5_checksum_fletcher32 (big-endian 16-bit words, odd tail byte in the high half)."""
    n = len(data)
    a = np.frombuffer(data[:n - (n % 2)], dtype=">u2").astype(np.uint64)
    s1 = s2 = 0
    # modular sums in blocks small enough not to overflow uint64
    pos = 0
    while pos < len(a):
        blk = a[pos:pos + 360]
        c = np.cumsum(blk)
        s2 = (s2 + s1 * len(blk) + int(c.sum())) % 65535
        s1 = (s1 + int(c[-1])) % 65535
        pos += 360
    if n % 2:
        s1 = (s1 + (data[-1] << 8)) % 65535
        s2 = (s2 + s1) % 65535
    return (s2 << 16) | s1


class _Writer:
    def __init__(self):
        self.buf = bytearray(96)   # superblock placeholder

    def alloc(self, data, align=8):
        self.buf += b"\0" * (-len(self.buf) % align)
        addr = len(self.buf)
        self.buf += data
        return addr

    def object_header(self, msgs):
        body = b""
        for mtype, mbody in msgs:
            mb = _pad8(mbody)
            body += struct.pack("<HHBBBB", mtype, len(mb), 0, 0, 0, 0) + mb
        hdr = struct.pack("<BBHII", 1, 0, len(msgs), 1, len(body)) + b"\0\0\0\0"
        return self.alloc(hdr + body)

    def write_dataset(self, d):
        a = _normalise(d.data)
        a = np.ascontiguousarray(a)
        msgs = []
        shape = a.shape
        if d.chunks:
            chunks = tuple(int(c) for c in d.chunks)
            msgs.append((0x1, _space_msg(shape, [None] + list(shape[1:]))))
        else:
            msgs.append((0x1, _space_msg(shape)))
        msgs.append((0x3, _dtype_msg(a.dtype)))
        msgs.append((0x5, struct.pack("<BBBB", 2, 2, 2, 0)))
        if d.chunks:
            rank = len(shape)
            es = a.dtype.itemsize
            if d.compress:
                flt = struct.pack("<BBHI", 1, 3, 0, 0)
                flt += struct.pack("<HHHH", 2, 0, 1, 1) + struct.pack("<II", es, 0)          # shuffle
                flt += struct.pack("<HHHH", 3, 0, 0, 0)                                        # fletcher32
                flt += struct.pack("<HHHH", 1, 0, 1, 1) + struct.pack("<II", 5, 0)            # deflate
                msgs.append((0xB, flt))
            recs = []
            grid = [range(0, s, c) for s, c in zip(shape, chunks)]
            import itertools
            for offs in itertools.product(*grid):
                ch = np.zeros(chunks, dtype=a.dtype)
                sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, chunks, shape))
                sub = tuple(slice(0, s.stop - s.start) for s in sl)
                ch[sub] = a[sl]
                raw = ch.tobytes()
                if d.compress:
                    ne = len(raw) // es
                    raw = np.frombuffer(raw, np.uint8).reshape(ne, es).T.tobytes()
                    raw = raw + struct.pack("<I", _fletcher32(raw))
                    raw = zlib.compress(raw, 5)
                recs.append((offs, len(raw), self.alloc(raw)))
            if len(recs) > 64:
                raise ValueError("writer supports at most 64 chunks per dataset")
            # single leaf B-tree node (type 1)
            node = b"TREE" + struct.pack("<BBHQQ", 1, 0, len(recs), UNDEF, UNDEF)
            for offs, csize, caddr in recs:
                node += struct.pack("<II", csize, 0) + struct.pack("<%dQ" % (rank + 1), *(list(offs) + [0]))
                node += struct.pack("<Q", caddr)
            last = [s for s in shape] + [0]
            node += struct.pack("<II", 0, 0) + struct.pack("<%dQ" % (rank + 1), *last)
            # pad to the full node size libhdf5 expects (2K=64 entries for chunk trees, K=32)
            full = 24 + 64 * (8 + 8 * (rank + 1) + 8) + 8 + 8 * (rank + 1)
            node += b"\0" * (full - len(node))
            baddr = self.alloc(node) if recs else UNDEF
            lay = struct.pack("<BBB", 3, 2, rank + 1) + struct.pack("<Q", baddr)
            lay += struct.pack("<%dI" % (rank + 1), *(list(chunks) + [es]))
            msgs.append((0x8, lay))
        else:
            raw = a.tobytes()
            addr = self.alloc(raw) if raw else UNDEF
            msgs.append((0x8, struct.pack("<BBQQ", 3, 1, addr, len(raw))))
        for k in sorted(d.attrs):
            msgs.append(_attr_msg(k, d.attrs[k]))
        return self.object_header(msgs)

    def write_group(self, g):
        names = sorted(g.children, key=lambda s: s.encode())
        addrs = {}
        for nm in names:
            c = g.children[nm]
            addrs[nm] = self.write_group(c) if isinstance(c, Group) else self.write_dataset(c)
        # local heap: offset 0 is the empty string
        heap = bytearray(b"\0" * 8)
        noff = {}
        for nm in names:
            noff[nm] = len(heap)
            heap += _pad8(nm.encode() + b"\0")
        free_off = len(heap)
        heap += struct.pack("<QQ", 1, 16)   # one free block: next=1 (none), size 16
        heap_data = self.alloc(bytes(heap))
        heap_addr = self.alloc(b"HEAP" + struct.pack("<BBBBQQQ", 0, 0, 0, 0, len(heap), free_off, heap_data))
        # symbol nodes: <= 2*K_leaf entries each (K_leaf written in the superblock)
        per = 2 * _K_LEAF
        snods = []
        for i in range(0, max(len(names), 1), per):
            part = names[i:i + per]
            sn = b"SNOD" + struct.pack("<BBH", 1, 0, len(part))
            for nm in part:
                sn += struct.pack("<QQII", noff[nm], addrs[nm], 0, 0) + b"\0" * 16
            sn += b"\0" * (40 * (per - len(part)))
            snods.append((self.alloc(sn), noff[part[-1]] if part else 0))
        if len(snods) > 2 * _K_INTERNAL:
            raise ValueError("too many children in one group")
        bt = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(snods), UNDEF, UNDEF)
        bt += struct.pack("<Q", 0)
        for saddr, lastoff in snods:
            bt += struct.pack("<QQ", saddr, lastoff)
        bt += b"\0" * (24 + 8 + 2 * _K_INTERNAL * 16 - len(bt))
        bt_addr = self.alloc(bt)
        msgs = [(0x11, struct.pack("<QQ", bt_addr, heap_addr))]
        for k in sorted(g.attrs):
            msgs.append(_attr_msg(k, g.attrs[k]))
        oh = self.object_header(msgs)
        g._bt, g._heap = bt_addr, heap_addr
        return oh


_K_LEAF = 4
_K_INTERNAL = 16


def save(root, path):
    """Serialise a ``Group`` tree to ``path`` (superblock v0, contiguous or chunked datasets)."""
    w = _Writer()
    oh = w.write_group(root)
    eof = len(w.buf)
    sb = SIGNATURE + struct.pack("<BBBBBBBB", 0, 0, 0, 0, 0, 8, 8, 0)
    sb += struct.pack("<HHI", _K_LEAF, _K_INTERNAL, 0)
    sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
    sb += struct.pack("<QQII", 0, oh, 1, 0) + struct.pack("<QQ", root._bt, root._heap)
    assert len(sb) == 96
    w.buf[:96] = sb
    with open(path, "wb") as fh:
        fh.write(bytes(w.buf))
