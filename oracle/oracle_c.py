"""TEST INFRASTRUCTURE -- ctypes binding of oracle/_build/liboracle.so (oracle.c).  PARITY UNPINNED.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    path = os.path.join(_HERE, "_build", "liboracle.so")
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return path


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        vp, u64p, u32p, u8p = C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_uint8)
        L.orc_db_build.restype = vp
        L.orc_db_build.argtypes = [u64p, C.c_uint32, C.c_uint32, C.c_uint32, u32p, C.c_uint32, C.c_int]
        L.orc_db_free.argtypes = [vp]
        L.orc_db_num_distinct.restype = C.c_uint64
        L.orc_db_num_distinct.argtypes = [vp]
        L.orc_db_num_entries.restype = C.c_uint64
        L.orc_db_num_entries.argtypes = [vp]
        L.orc_db_distinct_keys.argtypes = [vp, u64p]
        L.orc_query_begin.restype = vp
        L.orc_query_begin.argtypes = [vp, C.c_int, C.c_int, C.c_int]
        L.orc_query_free.argtypes = [vp]
        L.orc_query_push_ascii.argtypes = [vp, C.c_char_p, u64p, C.c_uint64]
        L.orc_query_push_packed.argtypes = [vp, u8p, u8p, u64p, C.c_uint64, C.c_uint32]
        L.orc_query_export_counts.argtypes = [vp, u8p]
        L.orc_query_import_counts.argtypes = [vp, u8p]
        L.orc_query_finish.argtypes = [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_double), u64p, u64p]
        L.orc_query_intersection.restype = C.c_uint64
        L.orc_query_intersection.argtypes = [vp, u64p, C.c_uint64]
        L.orc_max_threads.restype = C.c_int
        L.orc_set_threads.argtypes = [C.c_int]
        _LIB = L
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


GATE = {"exact": 0, "none": 1}


class OracleDB:
    """keys: uint64 array (G*n, 2) of (hi, lo); empty slot = (~0, ~0)."""

    def __init__(self, keys: np.ndarray, G: int, n: int, K: int = 60, ks=(30, 40, 50, 60), threads: int = 0):
        keys = np.ascontiguousarray(keys, dtype=np.uint64).reshape(-1)
        assert keys.size == 2 * G * n
        self.G, self.n, self.K, self.ks = G, n, K, tuple(int(k) for k in ks)
        ksa = np.asarray(self.ks, dtype=np.uint32)
        self._h = lib().orc_db_build(_p(keys, C.c_uint64), G, n, K, _p(ksa, C.c_uint32), len(self.ks), threads)
        if not self._h:
            raise ValueError("orc_db_build rejected its arguments")

    @property
    def num_distinct(self) -> int:
        return lib().orc_db_num_distinct(self._h)

    def distinct_keys(self) -> np.ndarray:
        out = np.empty((self.num_distinct, 2), dtype=np.uint64)
        lib().orc_db_distinct_keys(self._h, _p(out, C.c_uint64))
        return out

    def close(self):
        if self._h:
            lib().orc_db_free(self._h)
            self._h = None

    def __del__(self):
        self.close()


class OracleQuery:
    def __init__(self, db: OracleDB, ci_min: int = 2, gate: str = "exact", count_empty_in_den: bool = True):
        self.db = db
        self._h = lib().orc_query_begin(db._h, ci_min, GATE[gate], int(count_empty_in_den))
        if not self._h:
            raise ValueError("orc_query_begin rejected its arguments")

    def push_ascii(self, text: bytes, off: np.ndarray):
        off = np.ascontiguousarray(off, dtype=np.uint64)
        lib().orc_query_push_ascii(self._h, text, _p(off, C.c_uint64), off.size - 1)

    def push_reads(self, reads):
        text = "".join(reads).encode()
        off = np.zeros(len(reads) + 1, dtype=np.uint64)
        if len(reads):
            off[1:] = np.cumsum([len(r) for r in reads], dtype=np.uint64)
        self.push_ascii(text, off)

    def push_packed(self, bases: np.ndarray, nmask, off, nreads: int, read_len: int = 0):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        nm = None if nmask is None else np.ascontiguousarray(nmask, dtype=np.uint8)
        of = None if off is None else np.ascontiguousarray(off, dtype=np.uint64)
        lib().orc_query_push_packed(self._h, _p(bases, C.c_uint8), None if nm is None else _p(nm, C.c_uint8),
                                    None if of is None else _p(of, C.c_uint64), nreads, read_len)

    def export_counts(self) -> np.ndarray:
        out = np.empty(self.db.num_distinct, dtype=np.uint8)
        lib().orc_query_export_counts(self._h, _p(out, C.c_uint8))
        return out

    def import_counts(self, counts: np.ndarray):
        counts = np.ascontiguousarray(counts, dtype=np.uint8)
        assert counts.size == self.db.num_distinct
        lib().orc_query_import_counts(self._h, _p(counts, C.c_uint8))

    def export_sparse(self) -> np.ndarray:
        """the sparse form of export_counts(): int64 entries index | count << 32 for the non-zero clamped counters"""
        c = self.export_counts()
        idx = np.flatnonzero(c)
        return (idx.astype(np.int64) | (c[idx].astype(np.int64) << 32))

    def merge_sparse(self, own_dense: np.ndarray, others) -> None:
        """sum this rank's clamped counters and the other ranks' sparse entries, then import the sum"""
        tot = own_dense.astype(np.int64)
        for e in others:
            e = np.asarray(e, dtype=np.int64)
            e = e[(e >> 32) != 0]
            np.add.at(tot, e & 0xFFFFFFFF, e >> 32)
        self.import_counts(np.minimum(tot, 255).astype(np.uint8))

    def finish(self):
        G, nk = self.db.G, len(self.db.ks)
        num = np.zeros((G, nk), dtype=np.int64)
        den = np.zeros((G, nk), dtype=np.int64)
        ci = np.zeros((G, nk), dtype=np.float64)
        ni, nkm = C.c_uint64(0), C.c_uint64(0)
        lib().orc_query_finish(self._h, _p(num, C.c_int64), _p(den, C.c_int64), _p(ci, C.c_double),
                               C.byref(ni), C.byref(nkm))
        return dict(num=num, den=den, ci=ci, n_intersect=ni.value, n_kmers=nkm.value)

    def intersection(self) -> np.ndarray:
        n = lib().orc_query_intersection(self._h, None, 0)
        out = np.empty((n, 2), dtype=np.uint64)
        if n:
            lib().orc_query_intersection(self._h, _p(out, C.c_uint64), n)
        return out

    def close(self):
        if self._h:
            lib().orc_query_free(self._h)
            self._h = None

    def __del__(self):
        self.close()
