"""A small pure-Python HDF5 reader / writer for the files the reference exchanges.

The reference stores datasets (``series/mean``, ``series/max``, ``masks/raw`` ... + attr ``name``, datasets/nf.py:39-44,
113-125) and Keras models (``model_weights/<layer>/<weight>`` + JSON attributes, utils/keras_helpers.py:24-68,
unet_2d_summary.py:423-424,560) in HDF5 through h5py 2.7, i.e. in the library's CLASSIC on-disk format: superblock
version 0, version-1 object headers, groups as symbol tables (B-tree v1 + local heap), contiguous datasets,
fixed-length string attributes.  h5py / libhdf5 are not available in this environment, so this module implements
exactly that subset of the published HDF5 File Format Specification (v1.x / 2.0):

  reading : superblock 0 / 1 (user block allowed), symbol-table groups, object headers v1 with continuation blocks,
            dataspace v1 / v2, datatypes fixed-point / floating-point / fixed- and variable-length strings (global heap),
            data layouts compact / contiguous / chunked (layout messages v1-v3; deflate + shuffle filters),
            attribute messages v1-v3
  writing : superblock 0, symbol-table groups, contiguous datasets, numeric and fixed-length string attributes

Files with a version >= 2 superblock (h5py ``libver='latest'``) are rejected with a clear error.
The surface mirrors the part of h5py the reference uses: ``File(path, 'r'|'w')``, ``f['a/b']``, ``'a' in f``, ``keys()``,
``.attrs``, ``dataset[...]``, ``create_group``, ``create_dataset(name, data=...)``.

Pinned by tests/test_hdf5_lite.py against a genuine libhdf5-written file that ships with scipy
(scipy/io/matlab/tests/data/testhdf5_7.4_GLNX86.mat, a MATLAB v7.3 = HDF5 file with a 512-byte user block).
"""
import struct
import zlib

import numpy as np

SIGNATURE = b'\x89HDF\r\n\x1a\n'
UNDEF = 0xFFFFFFFFFFFFFFFF


class Hdf5Error(IOError):
    pass


# ====================================================================================================== reading
class _Reader(object):
    def __init__(self, buf):
        self.buf = buf
        self.base = 0
        off = 0
        while True:                                   # the superblock sits at 0, 512, 1024, ... (user block)
            if off + 8 > len(buf):
                raise Hdf5Error('not an HDF5 file (no superblock signature)')
            if buf[off:off + 8] == SIGNATURE:
                break
            off = 512 if off == 0 else off * 2
        ver = buf[off + 8]
        if ver > 1:
            raise Hdf5Error('HDF5 superblock version %d: only the classic format (versions 0 and 1, what h5py writes by '
                            'default with libver="earliest") is implemented' % ver)
        self.so, self.sl = buf[off + 13], buf[off + 14]          # size of offsets / lengths
        if (self.so, self.sl) != (8, 8):
            raise Hdf5Error('HDF5 offset / length size %d / %d not supported (8 / 8 only)' % (self.so, self.sl))
        p = off + 24 + (4 if ver == 1 else 0)
        self.base = self.u64(p)
        p += 32                                                   # base, free-space, end-of-file, driver-info addresses
        # root group symbol table entry
        self.root_header = self.u64(p + 8)

    # -- primitives (addresses in the file are relative to the base address)
    def u8(self, p):
        return self.buf[p]

    def u16(self, p):
        return struct.unpack_from('<H', self.buf, p)[0]

    def u32(self, p):
        return struct.unpack_from('<I', self.buf, p)[0]

    def u64(self, p):
        return struct.unpack_from('<Q', self.buf, p)[0]

    def addr(self, a):
        return a + self.base

    # -- object headers
    def messages(self, header_addr):
        """[(type, flags, data offset, size)] of a version-1 object header, following continuation messages"""
        p = self.addr(header_addr)
        if self.buf[p:p + 4] == b'OHDR':
            raise Hdf5Error('version-2 object headers are not implemented (classic format only)')
        if self.u8(p) != 1:
            raise Hdf5Error('object header version %d at 0x%x not supported' % (self.u8(p), p))
        nmsg, size = self.u16(p + 2), self.u32(p + 8)
        blocks = [(p + 16, size)]
        out = []
        while blocks and len(out) < nmsg:
            q, remaining = blocks.pop(0)
            end = q + remaining
            while q + 8 <= end and len(out) < nmsg:
                mtype, msize, mflags = self.u16(q), self.u16(q + 2), self.u8(q + 4)
                data = q + 8
                if mtype == 0x0010:                               # continuation: (offset, length)
                    blocks.append((self.addr(self.u64(data)), self.u64(data + 8)))
                out.append((mtype, mflags, data, msize))
                q = data + msize
        return out

    # -- groups: symbol table = B-tree v1 (type 0) over symbol-table nodes + local heap of names
    def heap_data(self, heap_addr):
        p = self.addr(heap_addr)
        if self.buf[p:p + 4] != b'HEAP':
            raise Hdf5Error('local heap signature missing at 0x%x' % p)
        return self.addr(self.u64(p + 24))

    def cstr(self, p):
        e = self.buf.index(b'\x00', p)
        return self.buf[p:e].decode('utf8')

    def group_entries(self, btree_addr, heap_addr):
        """{name: object header address} of a symbol-table group"""
        names = {}
        hd = self.heap_data(heap_addr)

        def walk(node):
            p = self.addr(node)
            sig = self.buf[p:p + 4]
            if sig == b'TREE':
                if self.u8(p + 4) != 0:
                    raise Hdf5Error('unexpected B-tree node type in a group')
                n = self.u16(p + 6)
                q = p + 24                                        # key 0
                for i in range(n):
                    walk(self.u64(q + 8))                         # child i follows key i
                    q += 16
            elif sig == b'SNOD':
                n = self.u16(p + 6)
                q = p + 8
                for i in range(n):
                    names[self.cstr(hd + self.u64(q))] = self.u64(q + 8)
                    q += 40
            else:
                raise Hdf5Error('unknown group node signature %r at 0x%x' % (sig, p))
        if btree_addr != UNDEF:
            walk(btree_addr)
        return names

    # -- datatypes
    def datatype(self, p):
        """-> (numpy dtype or ('vlen_str',) marker, bytes consumed are not needed)"""
        cv = self.u8(p)
        cls, b0, b1, size = cv & 0x0F, self.u8(p + 1), self.u8(p + 2), self.u32(p + 4)
        if cls == 0:                                              # fixed point
            order = '>' if (b0 & 1) else '<'
            return np.dtype('%s%s%d' % (order, 'i' if (b0 & 8) else 'u', size))
        if cls == 1:                                              # IEEE floating point
            order = '>' if (b0 & 1) else '<'
            if size not in (2, 4, 8):
                raise Hdf5Error('floating-point size %d not supported' % size)
            return np.dtype('%sf%d' % (order, size))
        if cls == 3:                                              # fixed-length string
            return np.dtype('S%d' % size)
        if cls == 9:                                              # variable length
            if (b0 & 0x0F) == 1:
                return 'vlen_str'
            raise Hdf5Error('variable-length sequences are not supported (strings only)')
        if cls == 8:                                              # enum (h5py booleans): read as the base integer type
            return self.datatype(p + 8)
        raise Hdf5Error('HDF5 datatype class %d not supported' % cls)

    def dataspace(self, p):
        ver = self.u8(p)
        rank = self.u8(p + 1)
        if ver == 1:
            q = p + 8
        elif ver == 2:
            if self.u8(p + 3) == 2:                               # null dataspace
                return None
            q = p + 4
        else:
            raise Hdf5Error('dataspace message version %d not supported' % ver)
        return tuple(self.u64(q + 8 * i) for i in range(rank))

    def global_heap_object(self, collection_addr, index):
        p = self.addr(collection_addr)
        if self.buf[p:p + 4] != b'GCOL':
            raise Hdf5Error('global heap signature missing at 0x%x' % p)
        end = p + self.u64(p + 8)
        q = p + 16
        while q + 16 <= end:
            idx, size = self.u16(q), self.u64(q + 8)
            if idx == 0:
                break
            if idx == index:
                return self.buf[q + 16:q + 16 + size]
            q += 16 + ((size + 7) // 8) * 8
        raise Hdf5Error('global heap object %d not found' % index)

    def decode(self, dtype, shape, raw):
        n = 1
        for v in (shape or ()):
            n *= v
        if isinstance(dtype, str):                                # variable-length strings: (length, heap address, index)
            out = []
            for i in range(n):
                ln, col, idx = struct.unpack_from('<IQI', raw, 16 * i)
                out.append(bytes(self.global_heap_object(col, idx)[:ln]).decode('utf8') if ln else '')
            a = np.array(out, dtype=object).reshape(shape or ())
            return a if shape else a[()]
        a = np.frombuffer(raw, dtype=dtype, count=n).reshape(shape or ())
        if shape:
            return a.copy()
        return a[()]

    # -- attributes
    def attribute(self, p):
        ver = self.u8(p)
        if ver == 1:
            nsz, tsz, ssz = self.u16(p + 2), self.u16(p + 4), self.u16(p + 6)
            pad = lambda v: (v + 7) // 8 * 8
            q = p + 8
            name = self.cstr(q); q += pad(nsz)
            tp = q; q += pad(tsz)
            sp = q; q += pad(ssz)
        elif ver in (2, 3):
            nsz, tsz, ssz = self.u16(p + 2), self.u16(p + 4), self.u16(p + 6)
            q = p + 8 + (1 if ver == 3 else 0)
            name = self.cstr(q); q += nsz
            tp = q; q += tsz
            sp = q; q += ssz
        else:
            raise Hdf5Error('attribute message version %d not supported' % ver)
        dtype = self.datatype(tp)
        shape = self.dataspace(sp)
        if shape is None:
            return name, None
        itemsize = 16 if isinstance(dtype, str) else dtype.itemsize
        n = 1
        for v in shape:
            n *= v
        return name, self.decode(dtype, shape, self.buf[q:q + n * itemsize])

    # -- dataset raw data
    def chunked(self, btree_addr, shape, chunk, itemsize, filters):
        rank = len(shape)
        out = np.zeros(shape, dtype='V%d' % itemsize)
        csize = itemsize
        for c in chunk:
            csize *= c

        def walk(node):
            p = self.addr(node)
            if self.buf[p:p + 4] != b'TREE' or self.u8(p + 4) != 1:
                raise Hdf5Error('chunk B-tree node expected at 0x%x' % p)
            level, n = self.u8(p + 5), self.u16(p + 6)
            keysz = 8 + 8 * (rank + 1)
            q = p + 24
            for i in range(n):
                nbytes, mask = self.u32(q), self.u32(q + 4)
                offs = [self.u64(q + 8 + 8 * d) for d in range(rank)]
                child = self.u64(q + keysz)
                if level > 0:
                    walk(child)
                else:
                    raw = bytes(self.buf[self.addr(child):self.addr(child) + nbytes])
                    for fid in reversed(filters):
                        if mask & (1 << filters.index(fid)):
                            continue
                        if fid == 1:
                            raw = zlib.decompress(raw)
                        elif fid == 2:                            # byte shuffle
                            a = np.frombuffer(raw, np.uint8).reshape(itemsize, -1)
                            raw = a.T.tobytes()
                        elif fid == 3:                            # fletcher32 checksum: strip it
                            raw = raw[:-4]
                        else:
                            raise Hdf5Error('HDF5 filter %d not supported' % fid)
                    block = np.frombuffer(raw[:csize], dtype='V%d' % itemsize).reshape(chunk)
                    sel = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, chunk, shape))
                    out[sel] = block[tuple(slice(0, s.stop - s.start) for s in sel)]
                q += keysz + 8
        if btree_addr != UNDEF:
            walk(btree_addr)
        return out.tobytes()


class _Attrs(dict):
    pass


class _Node(object):
    def __init__(self, rd, header_addr, name):
        self._rd, self._hdr, self.name = rd, header_addr, name
        self._msgs = rd.messages(header_addr)
        self.attrs = _Attrs()
        for mtype, mflags, p, size in self._msgs:
            if mtype == 0x000C:
                k, v = rd.attribute(p)
                self.attrs[k] = v


class Group(_Node):
    def __init__(self, rd, header_addr, name='/'):
        _Node.__init__(self, rd, header_addr, name)
        self._entries = {}
        for mtype, mflags, p, size in self._msgs:
            if mtype == 0x0011:
                self._entries = rd.group_entries(rd.u64(p), rd.u64(p + 8))
            elif mtype in (0x0002, 0x0006):
                raise Hdf5Error('new-style (link message) groups are not implemented (classic format only)')

    def keys(self):
        return sorted(self._entries.keys())

    def __iter__(self):
        return iter(self.keys())

    def __len__(self):
        return len(self._entries)

    def __contains__(self, path):
        try:
            self[path]
            return True
        except KeyError:
            return False

    def __getitem__(self, path):
        node = self
        for part in [s for s in path.split('/') if s]:
            if not isinstance(node, Group) or part not in node._entries:
                raise KeyError(path)
            node = node._open(part)
        return node

    def _open(self, part):
        hdr = self._entries[part]
        child_name = (self.name.rstrip('/') + '/' + part)
        types = {m[0] for m in self._rd.messages(hdr)}
        if 0x0011 in types or 0x0002 in types:
            return Group(self._rd, hdr, child_name)
        return Dataset(self._rd, hdr, child_name)

    def items(self):
        return [(k, self[k]) for k in self.keys()]


class Dataset(_Node):
    def __init__(self, rd, header_addr, name):
        _Node.__init__(self, rd, header_addr, name)
        self.dtype, self.shape, self._layout, self._filters = None, None, None, []
        for mtype, mflags, p, size in self._msgs:
            if mtype == 0x0003:
                self.dtype = rd.datatype(p)
            elif mtype == 0x0001:
                self.shape = rd.dataspace(p)
            elif mtype == 0x0008:
                self._layout = p
            elif mtype == 0x000B:                                 # filter pipeline
                ver, nf = rd.u8(p), rd.u8(p + 1)
                q = p + (8 if ver == 1 else 2)
                for i in range(nf):
                    fid = rd.u16(q)
                    if ver == 1:
                        nlen, ncd = rd.u16(q + 2), rd.u16(q + 6)
                        q += 8 + (nlen + 7) // 8 * 8 + 4 * ncd + (4 if ncd % 2 else 0)
                    elif fid >= 256:
                        nlen, ncd = rd.u16(q + 2), rd.u16(q + 6)
                        q += 8 + nlen + 4 * ncd
                    else:
                        ncd = rd.u16(q + 4)
                        q += 6 + 4 * ncd
                    self._filters.append(fid)
        if self.dtype is None or self._layout is None:
            raise Hdf5Error('%s: not a dataset (datatype / layout message missing)' % name)

    @property
    def ndim(self):
        return len(self.shape or ())

    def _raw(self):
        rd, p = self._rd, self._layout
        itemsize = 16 if isinstance(self.dtype, str) else self.dtype.itemsize
        n = itemsize
        for v in (self.shape or ()):
            n *= v
        ver = rd.u8(p)
        if ver in (1, 2):
            rank, cls = rd.u8(p + 1), rd.u8(p + 2)
            q = p + 8
            if cls == 0:                                          # compact: dims, size, data
                q += 4 * rank
                return rd.buf[q + 4:q + 4 + rd.u32(q)]
            a = rd.u64(q)
            dims = [rd.u32(q + 8 + 4 * i) for i in range(rank)]
            if cls == 1:
                return b'\x00' * n if a == UNDEF else rd.buf[rd.addr(a):rd.addr(a) + n]
            return rd.chunked(a, self.shape, tuple(dims[:-1]), itemsize, self._filters)
        if ver == 3:
            cls = rd.u8(p + 1)
            if cls == 0:
                return rd.buf[p + 4:p + 4 + rd.u16(p + 2)]
            if cls == 1:
                a = rd.u64(p + 2)
                return b'\x00' * n if a == UNDEF else rd.buf[rd.addr(a):rd.addr(a) + n]
            if cls == 2:
                rank = rd.u8(p + 2)
                a = rd.u64(p + 3)
                dims = [rd.u32(p + 11 + 4 * i) for i in range(rank)]
                return rd.chunked(a, self.shape, tuple(dims[:-1]), itemsize, self._filters)
        raise Hdf5Error('%s: data layout message version %d not supported' % (self.name, ver))

    def __getitem__(self, key):
        a = self._rd.decode(self.dtype, self.shape, bytes(self._raw()))
        if key is Ellipsis or key == ():
            return a
        return a[key]

    def __array__(self, dtype=None, copy=None):
        a = self[...]
        return a if dtype is None else np.asarray(a, dtype=dtype)

    @property
    def value(self):
        return self[...]


# ====================================================================================================== writing
def _pad8(b):
    return b + b'\x00' * (-len(b) % 8)


def _dtype_message(dt):
    dt = np.dtype(dt)
    if dt.kind == 'S':
        return struct.pack('<BBBBI', 0x13, 0x00, 0, 0, max(dt.itemsize, 1))          # class 3 v1, null-terminated ASCII
    if dt.kind in 'iu':
        bits = (8 if dt.kind == 'i' else 0) | (1 if dt.byteorder == '>' else 0)
        return struct.pack('<BBBBIHH', 0x10, bits, 0, 0, dt.itemsize, 0, dt.itemsize * 8)
    if dt.kind == 'f':
        order = 1 if dt.byteorder == '>' else 0
        if dt.itemsize == 4:
            props = struct.pack('<HHBBBBI', 0, 32, 23, 8, 0, 23, 127)
            sign = 31
        elif dt.itemsize == 8:
            props = struct.pack('<HHBBBBI', 0, 64, 52, 11, 0, 52, 1023)
            sign = 63
        elif dt.itemsize == 2:
            props = struct.pack('<HHBBBBI', 0, 16, 10, 5, 0, 10, 15)
            sign = 15
        else:
            raise Hdf5Error('cannot write floating-point size %d' % dt.itemsize)
        return struct.pack('<BBBBI', 0x11, 0x20 | order, sign, 0, dt.itemsize) + props    # mantissa normalisation: implied 1
    if dt.kind == 'b':
        return _dtype_message(np.uint8)
    raise Hdf5Error('cannot write numpy dtype %r' % (dt,))


def _dataspace_message(shape):
    if shape == ():
        return struct.pack('<BBBB4x', 1, 0, 0, 0)
    return struct.pack('<BBBB4x', 1, len(shape), 0, 0) + b''.join(struct.pack('<Q', int(v)) for v in shape)


def _as_attr_array(v):
    if isinstance(v, str):
        v = v.encode('utf8')
    if isinstance(v, bytes):
        return np.array(v, dtype='S%d' % max(len(v), 1))
    a = np.asarray(v)
    if a.dtype.kind == 'U':
        a = np.char.encode(a, 'utf8')
    if a.dtype.kind == 'O':
        a = np.array([x.encode('utf8') if isinstance(x, str) else x for x in a.ravel()]).reshape(a.shape)
    if a.dtype == np.bool_:
        a = a.astype(np.uint8)
    return a


def _message(mtype, data, flags=0):
    data = _pad8(data)
    return struct.pack('<HHB3x', mtype, len(data), flags) + data


def _attr_message(name, value):
    a = _as_attr_array(value)
    nm = name.encode('utf8') + b'\x00'
    tp, sp = _dtype_message(a.dtype), _dataspace_message(a.shape)
    body = struct.pack('<BBHHH', 1, 0, len(nm), len(tp), len(sp)) + _pad8(nm) + _pad8(tp) + _pad8(sp) + a.tobytes()
    if len(body) > 65000:
        raise Hdf5Error('attribute %r is too large for a version-1 object header (%d bytes)' % (name, len(body)))
    return _message(0x000C, body)


class _WGroup(object):
    def __init__(self, name):
        self.name, self.attrs, self.children = name, {}, {}

    def create_group(self, path):
        node = self
        for part in [s for s in path.split('/') if s]:
            if part not in node.children:
                node.children[part] = _WGroup(part)
            node = node.children[part]
            if not isinstance(node, _WGroup):
                raise Hdf5Error('%s is a dataset' % part)
        return node

    def require_group(self, path):
        return self.create_group(path)

    def create_dataset(self, path, shape=None, dtype=None, data=None):
        parts = [s for s in path.split('/') if s]
        parent = self.create_group('/'.join(parts[:-1])) if len(parts) > 1 else self
        if data is None:
            data = np.zeros(shape, dtype=dtype)
        a = np.asarray(data, dtype=dtype)
        a = np.ascontiguousarray(a) if a.ndim else a.copy()       # (ascontiguousarray would turn a scalar into shape (1,))
        ds = _WDataset(parts[-1], a)
        parent.children[parts[-1]] = ds
        return ds

    def __getitem__(self, path):
        node = self
        for part in [s for s in path.split('/') if s]:
            node = node.children[part]
        return node

    def __contains__(self, path):
        try:
            self[path]
            return True
        except (KeyError, AttributeError):
            return False

    def keys(self):
        return sorted(self.children)


class _WDataset(object):
    def __init__(self, name, data):
        self.name, self.data, self.attrs = name, data, {}
        self.shape, self.dtype = data.shape, data.dtype

    def __setitem__(self, key, value):
        self.data[key] = value

    def __getitem__(self, key):
        return self.data[key]


class _Writer(object):
    LEAF_K = 16            # up to 32 links per symbol-table node
    INTERNAL_K = 64        # up to 128 symbol-table nodes under the (single) B-tree node of a group: 4096 links

    def __init__(self):
        self.chunks = []   # (address, bytes)
        self.pos = 0

    def alloc(self, nbytes, align=8):
        self.pos = (self.pos + align - 1) // align * align
        a = self.pos
        self.pos += nbytes
        return a

    def put(self, addr, data):
        self.chunks.append((addr, data))

    def header(self, messages):
        body = b''.join(messages)
        hdr = struct.pack('<BBHII4x', 1, 0, len(messages), 1, len(body)) + body
        a = self.alloc(len(hdr))
        self.put(a, hdr)
        return a

    def dataset(self, ds):
        raw = ds.data.tobytes()
        da = self.alloc(max(len(raw), 1))
        self.put(da, raw)
        msgs = [_message(0x0001, _dataspace_message(ds.data.shape)),
                _message(0x0003, _dtype_message(ds.data.dtype), flags=1),
                _message(0x0005, struct.pack('<BBBBI', 1, 2, 2, 1, 0), flags=1),          # fill value v1: late alloc, undefined-size 0
                _message(0x0008, struct.pack('<BBQQ', 3, 1, da, len(raw)))]               # layout v3, contiguous
        msgs += [_attr_message(k, v) for k, v in ds.attrs.items()]
        return self.header(msgs)

    def group(self, g):
        names = sorted(g.children)
        child_hdr = {}
        for n in names:
            c = g.children[n]
            child_hdr[n] = self.group(c) if isinstance(c, _WGroup) else self.dataset(c)
        # local heap: offset 0 = empty string (8 bytes), then the names, 8-byte aligned each
        heap = bytearray(b'\x00' * 8)
        offs = {}
        for n in names:
            offs[n] = len(heap)
            heap += _pad8(n.encode('utf8') + b'\x00')
        free_off = len(heap)
        heap += struct.pack('<QQ', 1, 16)                         # one free block: (next = 1 = none, size)
        heap_data = self.alloc(len(heap))
        self.put(heap_data, bytes(heap))
        heap_hdr = self.alloc(32)
        self.put(heap_hdr, b'HEAP' + struct.pack('<B3xQQQ', 0, len(heap), free_off, heap_data))
        # symbol-table nodes of up to 2 * LEAF_K links each (sorted by name) under ONE leaf-level B-tree node
        per = 2 * self.LEAF_K
        groups = [names[i:i + per] for i in range(0, len(names), per)] or [[]]
        if len(groups) > 2 * self.INTERNAL_K:
            raise Hdf5Error('group %s has too many links (%d)' % (g.name, len(names)))
        snod_addrs = []
        for part in groups:
            snod = b'SNOD' + struct.pack('<BBH', 1, 0, len(part))
            for n in part:
                c = g.children[n]
                if isinstance(c, _WGroup):
                    snod += struct.pack('<QQII', offs[n], child_hdr[n], 1, 0) + struct.pack('<QQ', c._btree, c._heap)
                else:
                    snod += struct.pack('<QQII16x', offs[n], child_hdr[n], 0, 0)
            snod += b'\x00' * (8 + 40 * per - len(snod))
            a = self.alloc(len(snod))
            self.put(a, snod)
            snod_addrs.append(a)
        # B-tree node: key[0] = heap offset of "", key[i + 1] = heap offset of the largest name in child i
        nchild = len(groups) if names else 0
        btree = b'TREE' + struct.pack('<BBHQQ', 0, 0, nchild, UNDEF, UNDEF) + struct.pack('<Q', 0)
        for part, a in zip(groups, snod_addrs):
            if part:
                btree += struct.pack('<QQ', a, offs[part[-1]])
        btree += b'\x00' * (24 + (2 * self.INTERNAL_K + 1) * 8 + 2 * self.INTERNAL_K * 8 - len(btree))
        btree_addr = self.alloc(len(btree))
        self.put(btree_addr, btree)
        g._btree, g._heap = btree_addr, heap_hdr
        msgs = [_message(0x0011, struct.pack('<QQ', btree_addr, heap_hdr))]
        msgs += [_attr_message(k, v) for k, v in g.attrs.items()]
        return self.header(msgs)

    def write(self, root, path):
        self.pos = 96                                             # superblock v0 = 56 bytes + root entry 40 bytes
        root_hdr = self.group(root)
        eof = self.alloc(0)
        sb = SIGNATURE + struct.pack('<BBBBBBBBHHI', 0, 0, 0, 0, 0, 8, 8, 0, self.LEAF_K, self.INTERNAL_K, 0)
        sb += struct.pack('<QQQQ', 0, UNDEF, eof, UNDEF)
        sb += struct.pack('<QQII', 0, root_hdr, 1, 0) + struct.pack('<QQ', root._btree, root._heap)
        out = bytearray(eof)
        out[0:len(sb)] = sb
        for a, d in self.chunks:
            out[a:a + len(d)] = d
        with open(path, 'wb') as fp:
            fp.write(bytes(out))


# ====================================================================================================== File
class File(object):
    """``File(path, 'r')`` -> read-only Group surface; ``File(path, 'w')`` -> build a tree, written on close()."""

    def __init__(self, path, mode='r'):
        self.path, self.mode = path, mode
        if mode == 'r':
            with open(path, 'rb') as fp:
                self._rd = _Reader(fp.read())
            self._root = Group(self._rd, self._rd.root_header, '/')
        elif mode == 'w':
            self._root = _WGroup('/')
        else:
            raise ValueError("mode must be 'r' or 'w'")
        self.attrs = self._root.attrs

    def __getitem__(self, path):
        return self._root[path]

    def __contains__(self, path):
        return path in self._root

    def __iter__(self):
        return iter(self._root.keys())

    def keys(self):
        return self._root.keys()

    def create_group(self, path):
        return self._root.create_group(path)

    def require_group(self, path):
        return self._root.create_group(path)

    def create_dataset(self, path, shape=None, dtype=None, data=None):
        return self._root.create_dataset(path, shape, dtype, data)

    def close(self):
        if self.mode == 'w' and self._root is not None:
            _Writer().write(self._root, self.path)
        self._root = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        if exc[0] is None:
            self.close()
        return False
