"""Host-side encodings shared by the ctypes shim, the DB file format and the ingest code.

Conventions (the same ones the CUDA kernels use, see csrc/kmer.cuh):
  * base code A=0 C=1 G=2 T=3  (KMC's order; numeric order of a packed k-mer == lexicographic order)
  * a K-mer (K <= 63) is a 2K-bit integer, first base most significant, held as (hi, lo) uint64
  * an empty CMash sketch slot ('' in `CE._kmers`, local_tests/dump_kmers.py:10-14) is (~0, ~0)
  * packed read stream: base i -> byte i//4, bits 7-2*(i%4) .. 6-2*(i%4); N bases are packed as A
  * N mask: bit i -> byte i//8, bit 7-(i%8)  (numpy packbits 'big' order)
"""
from __future__ import annotations

import numpy as np

EMPTY = np.uint64(0xFFFFFFFFFFFFFFFF)

_CODE = np.full(256, 4, dtype=np.uint8)
for _i, _c in enumerate("ACGT"):
    _CODE[ord(_c)] = _i
    _CODE[ord(_c.lower())] = _i
_CHARS = np.frombuffer(b"ACGT", dtype=np.uint8)


def kmer_to_key(s: str):
    """ACGT string (len <= 63) -> (hi, lo); '' -> empty marker."""
    if s == "":
        return int(EMPTY), int(EMPTY)
    v = 0
    for ch in s:
        c = int(_CODE[ord(ch)])
        if c > 3:
            raise ValueError("non-ACGT character in sketch k-mer: %r" % ch)
        v = (v << 2) | c
    return (v >> 64) & 0xFFFFFFFFFFFFFFFF, v & 0xFFFFFFFFFFFFFFFF


def key_to_kmer(hi: int, lo: int, K: int) -> str:
    if int(hi) == int(EMPTY):
        return ""
    v = (int(hi) << 64) | int(lo)
    return "".join("ACGT"[(v >> (2 * (K - 1 - i))) & 3] for i in range(K))


def sketches_to_keys(sketches, K: int) -> np.ndarray:
    """list of G lists of n strings -> uint64 array (G*n, 2)."""
    G = len(sketches)
    n = len(sketches[0]) if G else 0
    out = np.empty((G * n, 2), dtype=np.uint64)
    for g, sk in enumerate(sketches):
        assert len(sk) == n, "every sketch must have the same number of slots"
        for j, s in enumerate(sk):
            assert s == "" or len(s) == K
            out[g * n + j] = kmer_to_key(s)
    return out


def keys_to_sketches(keys: np.ndarray, G: int, n: int, K: int):
    keys = np.asarray(keys, dtype=np.uint64).reshape(G * n, 2)
    return [[key_to_kmer(keys[g * n + j, 0], keys[g * n + j, 1], K) for j in range(n)] for g in range(G)]


def ascii_to_keys(buf: np.ndarray, K: int) -> np.ndarray:
    """uint8 array of shape (m, K) holding ACGT characters -> (m, 2) uint64 keys (vectorised)."""
    codes = _CODE[buf].astype(np.uint64)
    if (codes > 3).any():
        raise ValueError("non-ACGT character in sketch k-mer")
    m = buf.shape[0]
    hi = np.zeros(m, dtype=np.uint64)
    lo = np.zeros(m, dtype=np.uint64)
    for i in range(K):
        hi = (hi << np.uint64(2)) | (lo >> np.uint64(62))
        lo = (lo << np.uint64(2)) | codes[:, i]
    return np.stack([hi, lo], axis=1)


def ascii_slots_to_keys(buf: np.ndarray, K: int) -> np.ndarray:
    """like ascii_to_keys, but a slot that starts with NUL is an empty sketch slot ('' in CMash's CE._kmers): key (~0, ~0)"""
    buf = np.asarray(buf, dtype=np.uint8).reshape(-1, K)
    keys = np.full((buf.shape[0], 2), np.uint64(0xFFFFFFFFFFFFFFFF), dtype=np.uint64)
    full = buf[:, 0] != 0
    if full.any():
        keys[full] = ascii_to_keys(buf[full], K)
    return keys


def pack_reads(reads):
    """list of read strings -> (bases uint8[], nmask uint8[], off uint64[N+1]); buffers padded to 16 bytes."""
    lens = np.fromiter((len(r) for r in reads), dtype=np.uint64, count=len(reads))
    off = np.zeros(len(reads) + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    text = np.frombuffer("".join(reads).encode(), dtype=np.uint8)
    bases, nmask = pack_ascii(text)
    return bases, nmask, off


def pack_ascii(text: np.ndarray):
    """uint8 ASCII bases -> (bases, nmask) packed streams, each padded with zeros to a multiple of 16 bytes."""
    nb = int(text.size)
    codes = _CODE[text]
    isn = codes > 3
    c = np.where(isn, 0, codes).astype(np.uint8)
    pad4 = (-nb) % 4
    if pad4:
        c = np.concatenate([c, np.zeros(pad4, dtype=np.uint8)])
    c = c.reshape(-1, 4)
    b = (c[:, 0] << 6) | (c[:, 1] << 4) | (c[:, 2] << 2) | c[:, 3]
    m = np.packbits(isn.astype(np.uint8))
    return _pad16(b.astype(np.uint8)), _pad16(m)


def unpack_reads(bases: np.ndarray, nmask, off=None, nreads: int = 0, read_len: int = 0):
    """inverse of pack_reads -> list of strings (N where the mask is set)."""
    if off is None:
        off = np.arange(nreads + 1, dtype=np.uint64) * np.uint64(read_len)
    nb = int(off[-1])
    b = np.asarray(bases, dtype=np.uint8)
    codes = np.stack([(b >> 6) & 3, (b >> 4) & 3, (b >> 2) & 3, b & 3], axis=1).reshape(-1)[:nb]
    chars = _CHARS[codes].copy()
    if nmask is not None:
        isn = np.unpackbits(np.asarray(nmask, dtype=np.uint8))[:nb].astype(bool)
        chars[isn] = ord("N")
    s = chars.tobytes().decode()
    return [s[int(off[i]):int(off[i + 1])] for i in range(len(off) - 1)]


def nmask_to_runs(nmask: np.ndarray, nbases: int) -> np.ndarray:
    """packed N mask -> sorted (start, length) uint32 runs, shape (m, 2): the input of mlg_query_push_packed_nruns.
    Works on the non-zero mask bytes only (N is rare), so it is cheap on multi-gigabase streams."""
    m = np.asarray(nmask, dtype=np.uint8)[: (nbases + 7) // 8]
    nz = np.flatnonzero(m)
    if nz.size == 0:
        return np.zeros((0, 2), dtype=np.uint32)
    bits = np.unpackbits(m[nz]).reshape(-1, 8)
    pos = (nz[:, None] * 8 + np.arange(8)[None, :])[bits != 0]
    pos = pos[pos < nbases]
    if pos.size == 0:
        return np.zeros((0, 2), dtype=np.uint32)
    brk = np.flatnonzero(np.diff(pos) != 1)
    starts = np.concatenate([pos[:1], pos[brk + 1]])
    ends = np.concatenate([pos[brk], pos[-1:]]) + 1
    return np.ascontiguousarray(np.stack([starts, ends - starts], axis=1).astype(np.uint32))


def runs_to_nmask(runs: np.ndarray, nbases: int) -> np.ndarray:
    """inverse of nmask_to_runs (mask padded to 16 bytes)"""
    isn = np.zeros(nbases, dtype=np.uint8)
    for a, l in np.asarray(runs, dtype=np.int64).reshape(-1, 2):
        isn[a:a + l] = 1
    return _pad16(np.packbits(isn))


def _pad16(a: np.ndarray) -> np.ndarray:
    pad = (-a.size) % 16
    if pad or a.size == 0:
        a = np.concatenate([a, np.zeros(pad if a.size else 16, dtype=np.uint8)])
    return np.ascontiguousarray(a)


def _rev2_64(x: np.ndarray) -> np.ndarray:
    """reverse the order of the 32 two-bit groups of every uint64"""
    x = x.astype(np.uint64)
    for sh, m in ((2, 0x3333333333333333), (4, 0x0F0F0F0F0F0F0F0F), (8, 0x00FF00FF00FF00FF), (16, 0x0000FFFF0000FFFF)):
        m = np.uint64(m)
        x = ((x >> np.uint64(sh)) & m) | ((x & m) << np.uint64(sh))
    return (x >> np.uint64(32)) | (x << np.uint64(32))


def canonical_keys(keys: np.ndarray, K: int) -> np.ndarray:
    """(n, 2) uint64 (hi, lo) keys of K-mers -> their canonical forms (the smaller of k-mer and reverse complement),
    vectorised; same arithmetic as key_canon in csrc/kmer.cuh"""
    keys = np.ascontiguousarray(keys, dtype=np.uint64).reshape(-1, 2)
    hi, lo = keys[:, 0], keys[:, 1]
    fhi, flo = _rev2_64(~lo), _rev2_64(~hi)              # all 64 groups of the complemented value, reversed
    s = 128 - 2 * K                                      # drop the groups that came from the zero padding
    if s >= 64:
        rlo, rhi = fhi >> np.uint64(s - 64), np.zeros_like(fhi)
    elif s == 0:
        rlo, rhi = flo, fhi
    else:
        rlo = (flo >> np.uint64(s)) | (fhi << np.uint64(64 - s))
        rhi = fhi >> np.uint64(s)
    take_rc = (rhi < hi) | ((rhi == hi) & (rlo < lo))
    out = keys.copy()
    out[take_rc, 0] = rhi[take_rc]
    out[take_rc, 1] = rlo[take_rc]
    return out
