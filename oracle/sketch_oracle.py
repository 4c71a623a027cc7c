"""TEST INFRASTRUCTURE -- ctypes binding of oracle/_build/libsketch_oracle.so (sketch_oracle.c): CPU restatement of CMash's
bottom-n MinHash sketch construction.  PARITY UNPINNED beyond the hash function (see the C file's header).

Only tests/ and tests/bench_sketch.py's CPU leg import this."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
PRIME = 9999999999971


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "libsketch_oracle.so")
        src = os.path.join(_HERE, "sketch_oracle.c")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", _HERE, "-s"])
        L = C.CDLL(path)
        L.sko_murmur3_x64_128.argtypes = [C.c_char_p, C.c_uint64, C.c_uint32, C.POINTER(C.c_uint64 * 2)]
        L.sko_sketch_genomes.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
        _LIB = L
    return _LIB


def murmur3_x64_128(data: bytes, seed: int = 0):
    out = (C.c_uint64 * 2)()
    lib().sko_murmur3_x64_128(data, len(data), seed, C.byref(out))
    return out[0], out[1]


def sketch_genomes(genomes, n: int, K: int, prime: int = PRIME):
    """genomes: list of bytes/str (records of one genome joined by a non-ACGT byte) -> (mins [G,n] u64, counts [G,n] u32,
    kmers [G,n] of K-byte strings, b'' where unused)"""
    texts = [g.encode() if isinstance(g, str) else bytes(g) for g in genomes]
    off = np.zeros(len(texts) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(t) for t in texts])
    text = np.frombuffer(b"".join(texts) + b"N", dtype=np.uint8).copy()
    G = len(texts)
    mins = np.empty((G, n), dtype=np.uint64)
    counts = np.empty((G, n), dtype=np.uint32)
    kmers = np.zeros((G, n, K), dtype=np.uint8)
    lib().sko_sketch_genomes(text.ctypes.data, off.ctypes.data, G, n, K, prime, mins.ctypes.data, counts.ctypes.data, kmers.ctypes.data)
    return mins, counts, kmers
