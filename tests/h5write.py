"""TEST INFRASTRUCTURE: an independent, minimal HDF5 WRITER for the layout CMash's training database has
(`CountEstimators/<genome basename>/{mins,counts,kmers}` + attributes), in the flavour h5py writes by
default: superblock version 0, old-style groups (symbol-table message, version-1 B-tree, local heap), version-1 object
headers, contiguous datasets.  Used to test metalign_b200/h5min.py and scripts/make_db_from_h5.py where h5py is absent;
the reader is additionally pinned on a real HDF5 file (tests/test_h5min.py).  Written from the HDF5 File Format
Specification; shares no code with the reader."""
import struct

UNDEF = 0xFFFFFFFFFFFFFFFF
LEAF_K, INTERNAL_K = 4, 16


class _Out:
    def __init__(self, base):
        self.b = bytearray()
        self.base = base

    def alloc(self, n, align=8):
        while len(self.b) % align:
            self.b.append(0)
        a = len(self.b)
        self.b.extend(b"\0" * n)
        return a

    def put(self, a, data):
        self.b[a:a + len(data)] = data


def _msg(mtype, body, flags=0):
    body = body + b"\0" * (-len(body) % 8)
    return struct.pack("<HHB3x", mtype, len(body), flags) + body


def _ohdr(out, msgs):
    data = b"".join(msgs)
    a = out.alloc(16 + len(data))
    out.put(a, struct.pack("<BBHII4x", 1, 0, len(msgs), 1, len(data)) + data)
    return a


def _dataspace(shape):
    return struct.pack("<BBB5x", 1, len(shape), 0) + b"".join(struct.pack("<Q", d) for d in shape)


def _dtype_int(size, signed=True):
    return struct.pack("<BBBBI", 0x10, 0x08 if signed else 0, 0, 0, size) + struct.pack("<HH", 0, 8 * size)


def _dtype_str(size):
    return struct.pack("<BBBBI", 0x13, 0x01, 0, 0, size)        # null-padded ASCII


def _dataset(out, shape, dtype_msg, raw):
    d = out.alloc(len(raw))
    out.put(d, raw)
    layout = struct.pack("<BBQQ", 3, 1, d, len(raw))
    return _ohdr(out, [_msg(0x0001, _dataspace(shape)), _msg(0x0003, dtype_msg, flags=1), _msg(0x0008, layout)])


def _attr(name, dtype_msg, shape, raw):
    nm = name.encode() + b"\0"
    ds = _dataspace(shape)
    pad = lambda x: x + b"\0" * (-len(x) % 8)     # noqa: E731
    return _msg(0x000C, struct.pack("<BxHHH", 1, len(nm), len(dtype_msg), len(ds)) + pad(nm) + pad(dtype_msg) + pad(ds) + raw)


def _group(out, children, extra_msgs=()):
    """children: dict name -> object header address.  Returns the group's object header address."""
    names = sorted(children, key=lambda s: s.encode())
    # local heap: offset 0 = the empty string, then the names
    seg = bytearray(b"\0" * 8)
    noff = {}
    for nm in names:
        noff[nm] = len(seg)
        e = nm.encode() + b"\0"
        seg.extend(e + b"\0" * (-len(e) % 8))
    seg.extend(b"\0" * 16)                                    # a free block at the tail
    free_off = len(seg) - 16
    seg[free_off:free_off + 16] = struct.pack("<QQ", 1, 16)   # next free = 1 (none), size
    seg_a = out.alloc(len(seg))
    out.put(seg_a, bytes(seg))
    heap_a = out.alloc(32)
    out.put(heap_a, b"HEAP" + struct.pack("<B3xQQQ", 0, len(seg), free_off, seg_a))
    # symbol-table nodes of up to 2*LEAF_K entries
    level = []                                                # (address, heap offset of the largest name below)
    per = 2 * LEAF_K
    for i in range(0, max(1, len(names)), per):
        part = names[i:i + per]
        a = out.alloc(8 + per * 40)
        body = b"SNOD" + struct.pack("<BBH", 1, 0, len(part))
        for nm in part:
            body += struct.pack("<QQII16x", noff[nm], children[nm], 0, 0)
        out.put(a, body)
        level.append((a, noff[part[-1]] if part else 0))
    # B-tree levels of up to 2*INTERNAL_K children
    lvl = 0
    while True:
        nxt = []
        per_n = 2 * INTERNAL_K
        nodes = [level[i:i + per_n] for i in range(0, len(level), per_n)]
        addrs = [out.alloc(24 + 8 + per_n * 16) for _ in nodes]
        for j, (a, kids) in enumerate(zip(addrs, nodes)):
            body = b"TREE" + struct.pack("<BBHQQ", 0, lvl, len(kids), addrs[j - 1] if j else UNDEF,
                                         addrs[j + 1] if j + 1 < len(addrs) else UNDEF)
            body += struct.pack("<Q", 0)
            for ca, ko in kids:
                body += struct.pack("<QQ", ca, ko)
            out.put(a, body)
            nxt.append((a, kids[-1][1]))
        level = nxt
        lvl += 1
        if len(level) == 1:
            break
    btree_a = level[0][0]
    return _ohdr(out, [_msg(0x0011, struct.pack("<QQ", btree_a, heap_a))] + list(extra_msgs))


def write_cmash_h5(path, sketches, ksize, user_block=0):
    """sketches: dict genome name -> (mins: list of int, counts: list of int, kmers: list of bytes, '' for unused slots)"""
    out = _Out(user_block)
    sb = out.alloc(96)
    genome_groups = {}
    for name, (mins, counts, kmers) in sketches.items():
        n = len(kmers)
        d_mins = _dataset(out, (n,), _dtype_int(8), b"".join(struct.pack("<q", int(x)) for x in mins))
        d_counts = _dataset(out, (n,), _dtype_int(8), b"".join(struct.pack("<q", int(x)) for x in counts))
        d_kmers = _dataset(out, (n,), _dtype_str(ksize), b"".join(k.ljust(ksize, b"\0") for k in kmers))
        attrs = [_attr("class", _dtype_str(14), (), b"CountEstimator"), _attr("ksize", _dtype_int(8), (), struct.pack("<q", ksize)),
                 _attr("filename", _dtype_str(max(1, len(name))), (), name.encode())]
        genome_groups[name] = _group(out, {"mins": d_mins, "counts": d_counts, "kmers": d_kmers}, attrs)
    ce = _group(out, genome_groups)
    root = _group(out, {"CountEstimators": ce})
    eof = len(out.b)
    head = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, LEAF_K, INTERNAL_K, 0)
    head += struct.pack("<QQQQ", user_block, UNDEF, eof, UNDEF)     # addresses are relative to the base address (= the user block)
    head += struct.pack("<QQII16x", 0, root, 0, 0)
    out.put(sb, head)
    with open(path, "wb") as f:
        f.write(b"\0" * user_block)
        f.write(bytes(out.b))
