"""Native database file (.mlgdb): the flat replacement for the CMash training HDF5 the reference reads
at scripts/select_db.py:69 (group CountEstimators/<basename>/kmers).

Layout (little endian):
    8   magic "MLGDB001"
    4   K            sketch k-mer length
    4   n            sketch slots per genome
    8   G            genomes
    4   nk           number of queried prefix lengths
    32  ks[8]        prefix lengths, ascending (unused entries 0)
    8   names_bytes
    ..  names        '\n'-joined UTF-8 genome names (training-file basenames, sorted, = CMash's import order)
    ..  zero padding to a multiple of 16
    ..  keys         G*n entries of (hi u64, lo u64): stored-orientation K-mer, (~0, ~0) for an empty slot
The device structures (sorted arrays, bucket index, fingerprint table, prefix classes, denominators) are
rebuilt on the GPU at load time by mlg_db_load; the file holds only the source of truth.
"""
from __future__ import annotations

import struct

import numpy as np

MAGIC = b"MLGDB001"
MAGIC_BUILT = b"MLGDB002"     # same head; behind it the built device structures (csrc/dbfile.cu) instead of the keys
_HDR = struct.Struct("<8sIIQI8IQ")


def write(path: str, keys: np.ndarray, names, G: int, n: int, K: int, ks) -> None:
    keys = np.ascontiguousarray(keys, dtype="<u8").reshape(-1)
    if keys.size != 2 * G * n:
        raise ValueError("keys must hold G*n (hi, lo) pairs")
    if len(names) != G:
        raise ValueError("need one name per genome")
    ks = list(ks)
    blob = "\n".join(names).encode()
    hdr = _HDR.pack(MAGIC, K, n, G, len(ks), *(ks + [0] * (8 - len(ks))), len(blob))
    with open(path, "wb") as f:
        f.write(hdr)
        f.write(blob)
        pos = len(hdr) + len(blob)
        f.write(b"\0" * ((-pos) % 16))
        keys.tofile(f)


def read_header(path: str):
    with open(path, "rb") as f:
        raw = f.read(_HDR.size)
        if len(raw) != _HDR.size or raw[:8] not in (MAGIC, MAGIC_BUILT):
            raise ValueError("%s is not a .mlgdb file" % path)
        vals = _HDR.unpack(raw)
    _, K, n, G, nk = vals[:5]
    ks = list(vals[5:13])[:nk]
    return dict(K=K, n=n, G=G, ks=ks, names_bytes=vals[13], header_bytes=_HDR.size, built=(raw[:8] == MAGIC_BUILT))


def read_names(path: str):
    h = read_header(path)
    with open(path, "rb") as f:
        f.seek(h["header_bytes"])
        blob = f.read(h["names_bytes"])
    names = blob.decode().split("\n") if blob else []
    if len(names) != h["G"]:
        raise ValueError("%s: name table does not match G" % path)
    return names


def read_keys(path: str) -> np.ndarray:
    h = read_header(path)
    if h["built"]:
        raise ValueError("%s is a built database: it does not carry the source keys" % path)
    pos = h["header_bytes"] + h["names_bytes"]
    pos += (-pos) % 16
    return np.fromfile(path, dtype="<u8", offset=pos, count=2 * h["G"] * h["n"]).reshape(-1, 2)


def read_keys_rows(path: str, genomes) -> np.ndarray:
    """(m, n, 2) source keys of the listed genomes only (a seek per genome; source-form files)"""
    h = read_header(path)
    if h["built"]:
        raise ValueError("%s is a built database: it does not carry the source keys" % path)
    pos = h["header_bytes"] + h["names_bytes"]
    pos += (-pos) % 16
    n = h["n"]
    out = np.empty((len(genomes), n, 2), dtype=np.uint64)
    with open(path, "rb") as f:
        for i, g in enumerate(genomes):
            if not 0 <= int(g) < h["G"]:
                raise ValueError("genome %d out of range" % int(g))
            f.seek(pos + int(g) * n * 16)
            out[i] = np.frombuffer(f.read(n * 16), dtype="<u8").reshape(n, 2)
    return out


def keys_from_dump_fasta(path: str, K: int, expect_records: int = 0, chunk_bytes: int = 256 << 20) -> np.ndarray:
    """Keys of the FASTA dump that local_tests/dump_kmers.py:10-14 of the reference writes from the training HDF5:
    one '>seq<i>' header and one sequence line per sketch slot, in CountEstimator order; the sequence line of an
    unused slot is empty.  Vectorised and chunked (the default database has 2e8 records, ~14 GB of text)."""
    from . import codec
    out = []
    carry = b""
    pending_header = False          # a header line has been seen and its sequence line not yet
    with open(path, "rb") as f:
        while True:
            block = f.read(chunk_bytes)
            last = not block
            buf = carry + block
            if not buf:
                break
            if last and not buf.endswith(b"\n"):
                buf += b"\n"
            cut = buf.rfind(b"\n") + 1
            carry, buf = buf[cut:], buf[:cut]
            if buf:
                a = np.frombuffer(buf, dtype=np.uint8)
                nl = np.flatnonzero(a == 10)
                starts = np.concatenate([[0], nl[:-1] + 1])
                ends = nl.copy()
                cr = (ends > starts) & (a[np.maximum(ends - 1, 0)] == 13)
                ends = ends - cr
                is_hdr = (ends > starts) & (a[np.minimum(starts, a.size - 1)] == ord(">"))
                # a record = a header line followed by ONE line (possibly empty); anything else is malformed
                prev_hdr = np.concatenate([[pending_header], is_hdr[:-1]])
                if (is_hdr & prev_hdr).any() or (~is_hdr & ~prev_hdr & (ends > starts)).any():
                    raise ValueError("%s: expected alternating header / sequence lines" % path)
                seq = np.flatnonzero(prev_hdr & ~is_hdr)
                lens = ends[seq] - starts[seq]
                if ((lens != 0) & (lens != K)).any():
                    raise ValueError("%s: sketch k-mer of length other than %d" % (path, K))
                keys = np.full((seq.size, 2), codec.EMPTY, dtype=np.uint64)
                full = np.flatnonzero(lens == K)
                if full.size:
                    idx = starts[seq[full]][:, None] + np.arange(K)[None, :]
                    keys[full] = codec.ascii_to_keys(a[idx], K)
                out.append(keys)
                pending_header = bool(is_hdr[-1]) if is_hdr.size else pending_header
            if last:
                break
    if pending_header:                  # the file ended right after a header: its (empty) sequence line was cut off
        out.append(np.full((1, 2), codec.EMPTY, dtype=np.uint64))
    keys = np.concatenate(out) if out else np.zeros((0, 2), dtype=np.uint64)
    if expect_records and keys.shape[0] != expect_records:
        raise ValueError("%s: expected %d records, found %d" % (path, expect_records, keys.shape[0]))
    return keys
