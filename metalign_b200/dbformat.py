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
        if len(raw) != _HDR.size or raw[:8] != MAGIC:
            raise ValueError("%s is not a .mlgdb file" % path)
        vals = _HDR.unpack(raw)
    _, K, n, G, nk = vals[:5]
    ks = list(vals[5:13])[:nk]
    return dict(K=K, n=n, G=G, ks=ks, names_bytes=vals[13], header_bytes=_HDR.size)


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
    pos = h["header_bytes"] + h["names_bytes"]
    pos += (-pos) % 16
    return np.fromfile(path, dtype="<u8", offset=pos, count=2 * h["G"] * h["n"]).reshape(-1, 2)
