"""The 2K-bit k-mer arithmetic of csrc/kmer.cuh, compiled for the host and checked against strings."""
import ctypes as C
import os
import random
import subprocess

import pytest

from metalign_b200 import codec
from oracle import oracle_py

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def kh(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("kmer") / "libkmer_host.so")
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([gxx, "-O2", "-shared", "-fPIC", "-x", "c++", os.path.join(HERE, "kmer_host.cpp"), "-o", out])
    L = C.CDLL(out)
    u64 = C.c_ulonglong
    for f in (L.t_rc, L.t_canon):
        f.argtypes = [u64, u64, C.c_uint, C.POINTER(u64 * 2)]
    L.t_sub.argtypes = [u64, u64, C.c_uint, C.c_uint, C.c_uint, C.POINTER(u64 * 2)]
    L.t_prefix.argtypes = [u64, u64, C.c_uint, C.c_uint, C.POINTER(u64 * 2)]
    L.t_shl.argtypes = [u64, u64, C.c_uint, C.POINTER(u64 * 2)]
    L.t_hash.argtypes = [u64, u64, C.c_uint]
    L.t_hash.restype = u64
    L.t_bucket.argtypes = [u64, u64]
    L.t_bucket.restype = u64
    L.t_bucket.argtypes = [u64, C.c_uint]
    L.t_fword.argtypes = [u64, C.c_uint]
    L.t_fword.restype = C.c_uint
    L.t_fbit.argtypes = [u64]
    L.t_fbit.restype = C.c_uint
    L.t_fmask.argtypes = [u64, C.c_uint]
    L.t_fmask.restype = C.c_uint
    L.t_fp.argtypes = [u64]
    L.t_fp.restype = C.c_uint
    L.t_minimizer.argtypes = [u64, u64, C.c_uint]
    L.t_minimizer.restype = C.c_uint
    L.t_hash_sk.argtypes = [u64, u64, C.c_uint, C.c_uint]
    L.t_hash_sk.restype = u64
    L.t_mmer_mix.argtypes = [C.c_uint, C.c_uint]
    L.t_mmer_mix.restype = C.c_uint
    L.t_rev2_32.argtypes = [C.c_uint]
    L.t_rev2_32.restype = C.c_uint
    L.t_mz_order.argtypes = [C.c_uint, C.c_uint]
    L.t_mz_order.restype = C.c_uint
    L.t_mz_ident.argtypes = [C.c_uint, C.c_uint]
    L.t_mz_ident.restype = u64
    L.t_key_mz.argtypes = [u64, u64, C.c_uint, C.POINTER(u64 * 2)]
    L.t_mz_bit_index.argtypes = [u64, C.c_uint]
    L.t_mz_bit_index.restype = u64
    L.t_mz_bucket.argtypes = [u64, C.c_uint]
    L.t_mz_bucket.restype = C.c_uint
    L.t_mz_bit2.argtypes = [C.c_uint]
    L.t_mz_bit2.restype = C.c_uint
    return L


def _call(fn, s, *args):
    hi, lo = codec.kmer_to_key(s)
    out = (C.c_ulonglong * 2)()
    fn(hi, lo, *args, C.byref(out))
    return out[0], out[1]


def test_rc_canon_sub_prefix(kh):
    rng = random.Random(3)
    for _ in range(400):
        K = rng.randint(1, 63)
        s = "".join(rng.choice("ACGT") for _ in range(K))
        assert codec.key_to_kmer(*_call(kh.t_rc, s, K), K) == oracle_py.rc(s)
        assert codec.key_to_kmer(*_call(kh.t_canon, s, K), K) == oracle_py.canon(s)
        k = rng.randint(1, K)
        off = rng.randint(0, K - k)
        assert codec.key_to_kmer(*_call(kh.t_sub, s, K, off, k), k) == s[off:off + k]
        assert codec.key_to_kmer(*_call(kh.t_prefix, s, K, k), k) == s[:k]
        hi, lo = _call(kh.t_shl, s[:k], 2 * (K - k))
        assert codec.key_to_kmer(hi, lo, K) == s[:k] + "A" * (K - k)


def test_extremes(kh):
    for K in (1, 31, 32, 33, 60, 63):
        for ch in "ACGT":
            s = ch * K
            assert codec.key_to_kmer(*_call(kh.t_rc, s, K), K) == oracle_py.rc(s)


def test_hash_bucket_monotone_and_fp_nonzero(kh):
    rng = random.Random(5)
    hs = sorted(rng.getrandbits(64) for _ in range(2000))
    for bbits in (0, 1, 7, 20, 31):
        b = [kh.t_bucket(h, bbits) for h in hs]
        assert b == sorted(b) and max(b) < (1 << bbits)
    for nfw in (1, 1000, 10_000_000):
        assert all(kh.t_fword(h, nfw) < nfw for h in hs)
    assert all(kh.t_fbit(h) < 32 for h in hs)
    assert all(kh.t_fmask(h, 1) == 1 << kh.t_fbit(h) and bin(kh.t_fmask(h, 2)).count('1') in (1, 2) for h in hs)
    assert kh.t_fp(0) == 1 and kh.t_fp(1 << 31) == 1
    assert all(0 < kh.t_fp(h) < 2**31 for h in hs)
    # different keys hash differently (sanity, not a guarantee)
    assert len({kh.t_hash(rng.getrandbits(56), rng.getrandbits(64), 60) for _ in range(5000)}) == 5000
    # the hash is strand-symmetric: a k-mer and its reverse complement hash alike
    for _ in range(200):
        K = rng.randint(8, 63)
        s = "".join(rng.choice("ACGT") for _ in range(K))
        assert kh.t_hash(*codec.kmer_to_key(s), K) == kh.t_hash(*codec.kmer_to_key(oracle_py.rc(s)), K)


def _kmer_int(s):
    v = 0
    for ch in s:
        v = (v << 2) | "ACGT".index(ch)
    return v


def test_minimizer_is_strand_symmetric_and_matches_strings(kh):
    """super-k-mer layout: the minimizer of a K-mer is the smallest mixed value over its canonical 16-mers, the
    same for both strands; overlapping K-mers that share it agree on the top 32 bits of key_hash_sk (the bucket)."""
    rng = random.Random(5)
    for K in (60, 31, 16, 45):
        for _ in range(200):
            s = "".join(rng.choice("ACGT") for _ in range(K))
            r = oracle_py.rc(s)
            want = min(kh.t_mmer_mix(_kmer_int(s[p:p + 16]), _kmer_int(oracle_py.rc(s[p:p + 16]))) for p in range(K - 15))
            hi, lo = codec.kmer_to_key(s)
            rhi, rlo = codec.kmer_to_key(r)
            assert kh.t_minimizer(hi, lo, K) == want == kh.t_minimizer(rhi, rlo, K)
            for bbits in (2, 11, 27, 31):
                h = kh.t_hash_sk(hi, lo, K, bbits)
                assert h == kh.t_hash_sk(rhi, rlo, K, bbits)
                fp = kh.t_fp(h)
                assert fp == (h & 0x7FFFFFFF) and fp & 1
                # bucket = 2 * pair + half; the half is bit 13 of the fingerprint, the pair depends on the minimizer only
                assert kh.t_bucket(h, bbits) & 1 == (fp >> 13) & 1
    # canonical 16-mer mix: both argument orders agree, and rev2_32 reverses base order
    for _ in range(200):
        m = "".join(rng.choice("ACGT") for _ in range(16))
        f, r = _kmer_int(m), _kmer_int(oracle_py.rc(m))
        assert kh.t_mmer_mix(f, r) == kh.t_mmer_mix(r, f)
        assert kh.t_rev2_32(~f & 0xFFFFFFFF) == r
    # a long sequence: consecutive windows share minimizers in runs (super-k-mers), ~2/(w+1) changes per window
    seq = "".join(rng.choice("ACGT") for _ in range(5000))
    mins = []
    for i in range(len(seq) - 59):
        hi, lo = codec.kmer_to_key(seq[i:i + 60])
        mins.append(kh.t_minimizer(hi, lo, 60))
    changes = sum(1 for a, b in zip(mins, mins[1:]) if a != b)
    assert 0.02 < changes / len(mins) < 0.07


def _mz_pair(s32):
    """(a, b) of a 32-mer: its first 16 bases, and the reverse complement of its last 16 bases, as words"""
    return _kmer_int(s32[:16]), _kmer_int(oracle_py.rc(s32[16:]))


def _key_mz_strings(kh, s):
    """string-level restatement of key_mz: identities of the leftmost / rightmost 32-mer of smallest order"""
    vals = [kh.t_mz_order(*_mz_pair(s[p:p + 32])) for p in range(len(s) - 31)]
    m = min(vals)
    pl = vals.index(m)
    pr = len(vals) - 1 - vals[::-1].index(m)
    return kh.t_mz_ident(*_mz_pair(s[pl:pl + 32])), kh.t_mz_ident(*_mz_pair(s[pr:pr + 32]))


def test_minimizer32_layout_math(kh):
    """minimizer-bitmap layout (kmer.cuh): the order and the identity of a 32-mer do not depend on the strand; a K-mer's
    leftmost minimum is its reverse complement's rightmost one; low-complexity K-mers tie without changing identity"""
    rng = random.Random(9)
    out = (C.c_uint64 * 2)()
    for _ in range(300):
        m = "".join(rng.choice("ACGT") for _ in range(32))
        a, b = _mz_pair(m)
        ra, rb = _mz_pair(oracle_py.rc(m))
        assert (ra, rb) == (b, a)
        assert kh.t_mz_order(a, b) == kh.t_mz_order(b, a) < (1 << 26)
        assert kh.t_mz_ident(a, b) == kh.t_mz_ident(b, a)
    seqs = ["".join(rng.choice("ACGT") for _ in range(60)) for _ in range(300)]
    seqs += ["A" * 60, "AC" * 30, "ACG" * 20, "A" * 30 + "C" * 30, ("ACGTTGCA" * 8)[:60]]
    ties = 0
    for s in seqs:
        hi, lo = codec.kmer_to_key(s)
        kh.t_key_mz(hi, lo, 60, out)
        zl, zr = out[0], out[1]
        assert (zl, zr) == _key_mz_strings(kh, s)
        rhi, rlo = codec.kmer_to_key(oracle_py.rc(s))
        kh.t_key_mz(rhi, rlo, 60, out)
        assert (out[0], out[1]) == (zr, zl)
        ties += zl != zr
        for fbits in (5, 20, 32, 33, 36):
            assert kh.t_mz_bit_index(zl, fbits) == zl & ((1 << fbits) - 1)
        for bbits in (1, 13, 31):
            assert kh.t_mz_bucket(zl, bbits) == (zl >> 32) >> (32 - bbits)
        assert kh.t_mz_bit2(zl >> 32) == zl >> 59          # the second bit comes from identity bits outside the word index
    assert ties <= 2                      # homopolymers and short tandem repeats tie between 32-mers of EQUAL content
    # runs of equal minimizer along a sequence: ~2/(w+1) changes per window with w = 29 positions
    seq = "".join(rng.choice("ACGT") for _ in range(5000))
    ids = []
    for i in range(len(seq) - 59):
        hi, lo = codec.kmer_to_key(seq[i:i + 60])
        kh.t_key_mz(hi, lo, 60, out)
        ids.append(out[0])
    changes = sum(1 for x, y in zip(ids, ids[1:]) if x != y)
    assert 0.04 < changes / len(ids) < 0.10
