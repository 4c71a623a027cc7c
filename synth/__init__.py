"""Synthetic workload generator (TEST / BENCH INFRASTRUCTURE).

`SynthParams` mirrors `syn_params` in synth_core.h.  The CPU functions run anywhere; the
`cuda_*` functions need libmlg_synth_cuda.so and a GPU and write into device memory.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


class SynthParams(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("G", C.c_uint32), ("n", C.c_uint32), ("K", C.c_uint32),
                ("strain_period", C.c_uint32), ("tiny_pct", C.c_uint32), ("len_min", C.c_uint32),
                ("len_max", C.c_uint32), ("n_present", C.c_uint32), ("read_len", C.c_uint32),
                ("paired", C.c_uint32), ("sub_per_64k", C.c_uint32), ("n_per_64k", C.c_uint32)]


def params(G, n=1000, K=60, seed=20200529, strain_period=5, tiny_pct=1, len_min=1_000_000, len_max=5_000_000,
           n_present=500, read_len=150, paired=0, sub_per_64k=328, n_per_64k=66) -> SynthParams:
    assert len_min >= max(K, read_len * (2 if paired else 1)) and len_max >= len_min
    return SynthParams(seed, G, n, K, strain_period, tiny_pct, len_min, len_max, min(n_present, 1 << 20),
                       read_len, paired, sub_per_64k, n_per_64k)


_CPU = None
_CUDA = None


def _build(target, fname):
    path = os.path.join(_HERE, fname)
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".cu", ".h"))]
    if not os.path.exists(path) or os.path.getmtime(path) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s", target])
    return path


def cpu_lib():
    global _CPU
    if _CPU is None:
        L = C.CDLL(_build("libmlg_synth_cpu.so", "libmlg_synth_cpu.so"))
        pp = C.POINTER(SynthParams)
        L.syn_present_cum.argtypes = [pp, C.c_void_p]
        L.syn_gen_sketch_keys.argtypes = [pp, C.c_void_p]
        L.syn_gen_reads_ascii.argtypes = [pp, C.c_uint64, C.c_uint64, C.c_void_p]
        L.syn_gen_reads_packed.argtypes = [pp, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]
        L.syn_sizeof_params.restype = C.c_uint32
        assert L.syn_sizeof_params() == C.sizeof(SynthParams)
        _CPU = L
    return _CPU


def cuda_lib():
    """Device generator: same functions, device pointers (uintptr), plus a stream argument."""
    global _CUDA
    if _CUDA is None:
        L = C.CDLL(_build("libmlg_synth_cuda.so", "libmlg_synth_cuda.so"))
        pp = C.POINTER(SynthParams)
        L.syn_cuda_gen_sketch_keys.argtypes = [pp, C.c_void_p, C.c_void_p]
        L.syn_cuda_gen_sketch_keys.restype = C.c_int
        L.syn_cuda_gen_sketch_keys_range.argtypes = [pp, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
        L.syn_cuda_gen_sketch_keys_range.restype = C.c_int
        L.syn_cuda_gen_reads_packed.argtypes = [pp, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.syn_cuda_gen_reads_packed.restype = C.c_int
        _CUDA = L
    return _CUDA


def packed_sizes(nreads: int, read_len: int):
    """(bases bytes, nmask bytes) for a back-to-back packed stream, rounded up to 16 bytes."""
    nwords = (nreads * read_len + 63) // 64
    return ((nwords * 16 + 15) // 16) * 16, ((nwords * 8 + 15) // 16) * 16


def sketch_keys(p: SynthParams) -> np.ndarray:
    out = np.empty((p.G * p.n, 2), dtype=np.uint64)
    cpu_lib().syn_gen_sketch_keys(C.byref(p), out.ctypes.data)
    return out


def reads_ascii(p: SynthParams, r0: int, nreads: int) -> np.ndarray:
    out = np.empty((nreads, p.read_len), dtype=np.uint8)
    cpu_lib().syn_gen_reads_ascii(C.byref(p), r0, nreads, out.ctypes.data)
    return out


def reads_packed(p: SynthParams, r0: int, nreads: int):
    nbb, nmb = packed_sizes(nreads, p.read_len)
    bases = np.empty(nbb, dtype=np.uint8)
    nmask = np.empty(nmb, dtype=np.uint8)
    cpu_lib().syn_gen_reads_packed(C.byref(p), r0, nreads, bases.ctypes.data, nmask.ctypes.data)
    return bases, nmask
