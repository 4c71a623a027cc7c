"""Minimal reader for the HDF5 files CMash writes -- enough of the format to walk `CountEstimators/<genome>/kmers`
(`MinHash.export_multiple_to_single_hdf5`; read back by `import_multiple_from_single_hdf5`, which is what
`local_tests/dump_kmers.py:7-14` of the reference and CMash's query script walk) where h5py is not installed.

Supported (what h5py's default `libver='earliest'` produces): superblock version 0/1 (with or without a user block),
old-style groups (symbol-table message, version-1 B-trees, local heaps), version-1 object headers with continuation
blocks, simple dataspaces, fixed-point / floating-point / fixed-length string datatypes, compact, contiguous and chunked
layouts (chunk B-trees; deflate and shuffle filters), version-1/2/3 attributes of those types.  Anything else (version-2
object headers / new-style groups, variable-length strings, compound types) raises H5Unsupported with the reason.

Written from the HDF5 File Format Specification (version 1.1/2.0 structures named in the comments); no code from the
HDF5 library or h5py.  Pinned on a real file in the image -- scipy's `testhdf5_7.4_GLNX86.mat`, written by MATLAB's HDF5
library -- and on files made by an independent writer (tests/h5write.py); see tests/test_h5min.py.
"""
from __future__ import annotations

import mmap
import struct
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Unsupported(ValueError):
    pass


def _guard(fn):
    """damaged input (offsets beyond the file, impossible sizes) surfaces as H5Unsupported, never as struct.error / IndexError"""
    import functools

    @functools.wraps(fn)
    def wrapped(*a, **k):
        try:
            return fn(*a, **k)
        except (struct.error, IndexError, OverflowError, MemoryError, TypeError, zlib.error) as e:
            raise H5Unsupported("damaged or unsupported HDF5 structure: %s" % e) from None
        except ValueError as e:
            if isinstance(e, H5Unsupported):
                raise
            raise H5Unsupported("damaged or unsupported HDF5 structure: %s" % e) from None
    return wrapped


class H5File:
    @_guard
    def __init__(self, path: str):
        self._f = open(path, "rb")
        try:
            self.buf = mmap.mmap(self._f.fileno(), 0, access=mmap.ACCESS_READ)
        except ValueError:
            self._f.close()
            raise H5Unsupported("%s: empty file" % path)
        self.path = path
        b = self.buf
        sig = b"\x89HDF\r\n\x1a\n"
        off = 0
        while True:                                   # the superblock sits at 0, 512, 1024, ... (after a user block)
            if off + 8 > len(b):
                raise H5Unsupported("%s: no HDF5 signature" % path)
            if b[off:off + 8] == sig:
                break
            off = 512 if off == 0 else off * 2
        ver = b[off + 8]
        if ver > 1:
            raise H5Unsupported("%s: superblock version %d (written with libver='latest'?); only 0 and 1 are read" % (path, ver))
        self.so, self.sl = b[off + 13], b[off + 14]
        if self.so != 8 or self.sl != 8:
            raise H5Unsupported("%s: %d-byte offsets / %d-byte lengths" % (path, self.so, self.sl))
        p = off + 24 + (4 if ver == 1 else 0)
        self.base, _, self.eof, _ = struct.unpack_from("<QQQQ", b, p)
        root = p + 32                                  # root group symbol table entry
        _, ohdr, cache, _ = struct.unpack_from("<QQII", b, root)
        self.root = Group(self, ohdr, "/")

    def close(self):
        self.buf.close()
        self._f.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __getitem__(self, name):
        return self.root[name]

    def keys(self):
        return self.root.keys()

    # ---- raw access (file addresses are relative to the base address)
    def at(self, addr: int) -> int:
        if addr == UNDEF:
            raise H5Unsupported("undefined address")
        a = addr + self.base
        if a >= len(self.buf):
            raise H5Unsupported("%s: address %d beyond the end of the file (truncated?)" % (self.path, addr))
        return a

    @_guard
    def messages(self, ohdr_addr: int):
        """(type, flags, payload bytes) of every message of a version-1 object header, continuation blocks included"""
        b = self.buf
        a = self.at(ohdr_addr)
        if b[a:a + 4] == b"OHDR":
            raise H5Unsupported("version-2 object header (file written with libver='latest'); only version 1 is read")
        if b[a] != 1:
            raise H5Unsupported("object header version %d" % b[a])
        nmsg, = struct.unpack_from("<H", b, a + 2)
        size, = struct.unpack_from("<I", b, a + 8)
        blocks = [(a + 16, size)]
        out = []
        nblocks = 0
        while blocks and len(out) < nmsg:
            nblocks += 1
            if nblocks > 4096:
                raise H5Unsupported("object header continuation chain too long (damaged file)")
            p, left = blocks.pop(0)
            end = p + left
            while p + 8 <= end and len(out) < nmsg:
                mtype, msize, mflags = struct.unpack_from("<HHB", b, p)
                body = bytes(b[p + 8:p + 8 + msize])
                p += 8 + msize
                if mtype == 0x0010:                    # object header continuation: offset, length
                    co, cl = struct.unpack_from("<QQ", body, 0)
                    blocks.append((self.at(co), cl))
                out.append((mtype, mflags, body))
        return out


def _parse_dataspace(body: bytes):
    ver, rank, flags = body[0], body[1], body[2]
    if ver == 1:
        p = 8
    elif ver == 2:
        if body[3] == 2:                               # null dataspace
            return None
        p = 4
    else:
        raise H5Unsupported("dataspace version %d" % ver)
    return tuple(struct.unpack_from("<%dQ" % rank, body, p)) if rank else ()


def _parse_datatype(body: bytes):
    """numpy dtype of a datatype message (fixed-point, floating-point, fixed-length string)"""
    cls, bits0 = body[0] & 15, body[1]
    size, = struct.unpack_from("<I", body, 4)
    if size == 0 or size > (1 << 20):
        raise H5Unsupported("datatype of %d bytes" % size)
    if cls == 0:
        order = ">" if bits0 & 1 else "<"
        return np.dtype("%s%s%d" % (order, "i" if bits0 & 8 else "u", size))
    if cls == 1:
        order = ">" if bits0 & 1 else "<"
        return np.dtype("%sf%d" % (order, size))
    if cls == 3:
        return np.dtype("S%d" % size)
    if cls == 9:
        raise H5Unsupported("variable-length datatype (h5py str / vlen); CMash stores fixed-length byte strings")
    raise H5Unsupported("datatype class %d" % cls)


class _Object:
    def __init__(self, f: H5File, addr: int, name: str):
        self.file, self.addr, self.name = f, addr, name
        self._msgs = None

    @property
    def msgs(self):
        if self._msgs is None:
            self._msgs = self.file.messages(self.addr)
        return self._msgs

    @property
    def attrs(self) -> dict:
        return self._attrs()

    @_guard
    def _attrs(self) -> dict:
        out = {}
        for t, _, body in self.msgs:
            if t != 0x000C:
                continue
            ver = body[0]
            nlen, tlen, slen = struct.unpack_from("<HHH", body, 2)
            if ver == 1:
                pad = lambda x: (x + 7) & ~7        # noqa: E731
                p = 8
            elif ver in (2, 3):
                pad = lambda x: x                   # noqa: E731
                p = 8 + (1 if ver == 3 else 0)
            else:
                continue
            name = body[p:p + nlen].split(b"\0")[0].decode("utf-8", "replace")
            p += pad(nlen)
            try:
                dt = _parse_datatype(body[p:p + tlen])
            except H5Unsupported:
                continue
            p += pad(tlen)
            shape = _parse_dataspace(body[p:p + slen])
            p += pad(slen)
            if shape is None:
                out[name] = None
                continue
            n = int(np.prod(shape)) if shape else 1
            val = np.frombuffer(body, dtype=dt, count=n, offset=p)
            out[name] = val.reshape(shape) if shape else val[0]
        return out


class Group(_Object):
    def read(self):
        raise H5Unsupported("%s is a group, not a dataset" % self.name)

    def _table(self):
        for t, _, body in self.msgs:
            if t == 0x0011:
                return struct.unpack_from("<QQ", body, 0)
            if t in (0x0002, 0x0006):
                raise H5Unsupported("new-style group (link messages); CMash files written by h5py's default settings use symbol tables")
        raise H5Unsupported("%s is not a group" % self.name)

    @_guard
    def _entries(self):
        """name -> object header address, from the group's B-tree of symbol-table nodes and its local heap"""
        f, b = self.file, self.file.buf
        btree, heap = self._table()
        h = f.at(heap)
        if b[h:h + 4] != b"HEAP":
            raise H5Unsupported("local heap signature missing")
        data = f.at(struct.unpack_from("<Q", b, h + 24)[0])
        out = {}
        stack = [btree]
        seen = set()
        while stack:
            node = stack.pop()
            if node in seen or len(seen) > (1 << 24):
                raise H5Unsupported("group B-tree revisits a node (damaged file)")
            seen.add(node)
            a = f.at(node)
            if b[a:a + 4] == b"TREE":
                ntype, level, used = struct.unpack_from("<BBH", b, a + 4)
                if ntype != 0:
                    raise H5Unsupported("group B-tree of node type %d" % ntype)
                p = a + 24                              # key 0, then (child, key) pairs
                for i in range(used):
                    child, = struct.unpack_from("<Q", b, p + 8 + i * 16)
                    stack.append(child)
            elif b[a:a + 4] == b"SNOD":
                n, = struct.unpack_from("<H", b, a + 6)
                for i in range(n):
                    noff, oaddr = struct.unpack_from("<QQ", b, a + 8 + i * 40)
                    s = data + noff
                    e = b.find(b"\0", s)
                    out[bytes(b[s:e]).decode("utf-8", "replace")] = oaddr
            else:
                raise H5Unsupported("neither a B-tree node nor a symbol-table node at %d" % a)
        return out

    def keys(self):
        if getattr(self, "_names", None) is None:
            self._names = self._entries()
        return sorted(self._names)

    def __contains__(self, name):
        self.keys()
        return name in self._names

    @_guard
    def __getitem__(self, name: str):
        node = self
        for part in [x for x in name.split("/") if x]:
            node.keys()
            if part not in node._names:
                raise KeyError(part)
            addr = node._names[part]
            child_name = node.name.rstrip("/") + "/" + part
            types = {t for t, _, _ in node.file.messages(addr)}
            node = Dataset(node.file, addr, child_name) if 0x0008 in types else Group(node.file, addr, child_name)
        return node


class Dataset(_Object):
    def keys(self):
        raise H5Unsupported("%s is a dataset, not a group" % self.name)

    def __getitem__(self, name):
        raise H5Unsupported("%s is a dataset, not a group" % self.name)

    @_guard
    def _meta(self):
        shape = dtype = layout = None
        filters = []
        for t, _, body in self.msgs:
            if t == 0x0001:
                shape = _parse_dataspace(body)
            elif t == 0x0003:
                dtype = _parse_datatype(body)
            elif t == 0x0008:
                layout = body
            elif t == 0x000B:
                ver, nf = body[0], body[1]
                p = 8 if ver == 1 else 2
                for _ in range(nf):
                    fid, nlen, fl, ncd = struct.unpack_from("<HHHH", body, p)
                    if ver == 2 and fid < 256:
                        nlen = 0
                        fid, fl, ncd = struct.unpack_from("<HHH", body, p)
                        p += 6
                    else:
                        p += 8
                    p += (nlen + 7) & ~7 if ver == 1 else nlen
                    cd = struct.unpack_from("<%dI" % ncd, body, p)
                    p += 4 * ncd
                    if ver == 1 and ncd % 2:
                        p += 4
                    filters.append((fid, cd))
        if shape is None or dtype is None or layout is None:
            raise H5Unsupported("%s: dataspace, datatype or layout message missing" % self.name)
        return shape, dtype, layout, filters

    @property
    def shape(self):
        return self._meta()[0]

    @property
    def dtype(self):
        return self._meta()[1]

    @_guard
    def read(self) -> np.ndarray:
        f, b = self.file, self.file.buf
        shape, dtype, lay, filters = self._meta()
        n = 1
        for d in shape:
            n *= int(d)
        if n * dtype.itemsize > (1 << 40):
            raise H5Unsupported("%s: dataset of %d elements" % (self.name, n))
        ver = lay[0]
        if ver == 3:
            cls = lay[1]
            if cls == 0:                               # compact
                size, = struct.unpack_from("<H", lay, 2)
                return np.frombuffer(lay, dtype=dtype, count=n, offset=4).reshape(shape).copy()
            if cls == 1:                               # contiguous
                addr, size = struct.unpack_from("<QQ", lay, 2)
                if addr == UNDEF:                      # never written: fill value (zeros)
                    return np.zeros(shape, dtype=dtype)
                a = f.at(addr)
                if a + n * dtype.itemsize > len(b):
                    raise H5Unsupported("%s: data beyond the end of the file (truncated?)" % self.name)
                return np.frombuffer(b, dtype=dtype, count=n, offset=a).reshape(shape).copy()
            if cls == 2:                               # chunked
                rank = lay[2]
                btree, = struct.unpack_from("<Q", lay, 3)
                cdims = struct.unpack_from("<%dI" % rank, lay, 11)
                return self._read_chunks(shape, dtype, btree, cdims[:-1], filters)
            raise H5Unsupported("layout class %d" % cls)
        if ver in (1, 2):
            rank, cls = lay[1], lay[2]
            p = 8
            addr = UNDEF
            if cls != 0:
                addr, = struct.unpack_from("<Q", lay, p)
                p += 8
            dims = struct.unpack_from("<%dI" % rank, lay, p)
            if cls == 1:
                a = f.at(addr)
                return np.frombuffer(b, dtype=dtype, count=n, offset=a).reshape(shape).copy()
            if cls == 2:
                return self._read_chunks(shape, dtype, addr, dims[:-1], filters)
            p += 4 * rank
            size, = struct.unpack_from("<I", lay, p)
            return np.frombuffer(lay, dtype=dtype, count=n, offset=p + 4).reshape(shape).copy()
        raise H5Unsupported("data layout message version %d" % ver)

    def _read_chunks(self, shape, dtype, btree, cdims, filters):
        f, b = self.file, self.file.buf
        out = np.zeros(shape, dtype=dtype)
        if btree == UNDEF:
            return out
        rank = len(shape)
        stack = [btree]
        seen = set()
        while stack:
            node = stack.pop()
            if node in seen:
                raise H5Unsupported("chunk B-tree revisits a node (damaged file)")
            seen.add(node)
            a = f.at(node)
            if b[a:a + 4] != b"TREE":
                raise H5Unsupported("chunk B-tree node signature missing")
            ntype, level, used = struct.unpack_from("<BBH", b, a + 4)
            if ntype != 1:
                raise H5Unsupported("chunk B-tree of node type %d" % ntype)
            ksz = 8 + 8 * (rank + 1)
            p = a + 24
            for i in range(used):
                csize, fmask = struct.unpack_from("<II", b, p)
                offs = struct.unpack_from("<%dQ" % rank, b, p + 8)
                child, = struct.unpack_from("<Q", b, p + ksz)
                p += ksz + 8
                if level > 0:
                    stack.append(child)
                    continue
                ca = f.at(child)
                raw = bytes(b[ca:ca + csize])
                for k, (fid, cd) in reversed(list(enumerate(filters))):
                    if fmask & (1 << k):
                        continue
                    if fid == 1:
                        raw = zlib.decompress(raw)
                    elif fid == 2:                      # shuffle: bytes of every element de-interleaved
                        es = cd[0] if cd else dtype.itemsize
                        arr = np.frombuffer(raw, dtype=np.uint8)
                        m = arr.size // es
                        raw = arr[: m * es].reshape(es, m).T.tobytes() + arr[m * es:].tobytes()
                    elif fid == 3:                      # fletcher32 checksum appended: drop it
                        raw = raw[:-4]
                    else:
                        raise H5Unsupported("filter %d" % fid)
                chunk = np.frombuffer(raw, dtype=dtype, count=int(np.prod(cdims))).reshape(cdims)
                sl_out = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cdims, shape))
                sl_in = tuple(slice(0, s.stop - s.start) for s in sl_out)
                out[sl_out] = chunk[sl_in]
        return out
