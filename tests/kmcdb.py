"""TEST INFRASTRUCTURE -- an independent, pure-Python encoder / decoder of KMC k-mer database files
(`<name>.kmc_pre` + `<name>.kmc_suf`), the on-disk form of the artefacts the reference exchanges between its
subprocesses: `data/cmash_db_n1000_k60_dump` (scripts/select_db.py:44, written by `kmc -k60 -fa -ci0 -cs3`,
local_tests/retrain_and_test_metalign.sh:66), `reads_60mers` (select_db.py:50-52) and `60mers_intersection`
(select_db.py:54-56).

KMC is not vendored under /root/reference and is not installed here, so the layout below is the published one
(KMC API documentation, "k-mer database format") written down from memory -- UNPINNED against files
made by a real `kmc`.  Both variants are produced: version 0 ("KMC1": one prefix table; what `kmc_tools` writes)
and version 0x200 ("KMC2": what `kmc` itself writes -- one prefix table per signature bin plus a signature map).

    .kmc_pre   "KMCP" | uint64 LUT[...] (+ one guard entry = total) | [uint32 signature map (0x200 only)] | header |
               uint32 header_offset | "KMCP"
               LUT entry i = number of records in .kmc_suf before the records of prefix (i mod 4^lut_prefix_length)
               of table (i div 4^lut_prefix_length); header = kmer_length, mode, counter_size, lut_prefix_length,
               [signature_len], min_count, max_count, total_kmers (u64), !both_strands (u8), 3 pad bytes,
               max_count high word, 20 reserved bytes, kmc_version; header_offset = bytes of the header
    .kmc_suf   "KMCS" | records: (k - lut_prefix_length)/4 bytes of suffix bases, first base in the top bits,
               then counter_size bytes of count, little endian | "KMCS"

Used by the stub `kmc` / `kmc_tools` / `kmc_dump` executables under tests/golden/stub_tools/ and by the tests of
the product's C++ reader (metalign_b200/csrc/kmcdb.cpp).
"""
from __future__ import annotations

import struct
from typing import Dict, Iterable, List, Tuple

_CODE = {"A": 0, "C": 1, "G": 2, "T": 3}
_BASE = "ACGT"


def _to_int(kmer: str) -> int:
    v = 0
    for ch in kmer:
        v = (v << 2) | _CODE[ch]
    return v


def _to_str(v: int, k: int) -> str:
    return "".join(_BASE[(v >> (2 * (k - 1 - i))) & 3] for i in range(k))


def choose_lut_prefix_length(k: int, want: int = 0) -> int:
    """KMC keeps (k - lut_prefix_length) a multiple of 4 so that suffixes are whole bytes"""
    p = want if want else max(1, min(k - 4, 4))
    while (k - p) % 4:
        p += 1
    if p >= k:
        raise ValueError("k too small for a byte-aligned suffix")
    return p


def write(prefix: str, kmers: Dict[str, int], k: int, counter_size: int = 1, min_count: int = 1, max_count: int = 255,
          canonical: bool = True, version: int = 0, lut_prefix_length: int = 0, signature_len: int = 5, n_bins: int = 3) -> None:
    """kmers: k-mer string -> count.  version 0 or 0x200."""
    if version not in (0, 0x200):
        raise ValueError("kmc_version must be 0 or 0x200")
    p = choose_lut_prefix_length(k, lut_prefix_length)
    suf_bytes = (k - p) // 4
    suf_bits = 2 * (k - p)
    cmask = (1 << (8 * counter_size)) - 1
    items: List[Tuple[int, int]] = sorted((_to_int(s), c) for s, c in kmers.items())
    if version == 0:
        bins = [items]
    else:
        # KMC2 files group records by signature bin; any assignment of k-mers to bins is a valid file as long as
        # every bin is sorted -- here: by a cheap function of the k-mer value
        bins = [[] for _ in range(n_bins)]
        for v, c in items:
            bins[(v * 2654435761 >> 7) % n_bins].append((v, c))
    lut: List[int] = []
    suf = bytearray(b"KMCS")
    n_rec = 0
    for recs in bins:
        table = [0] * (1 << (2 * p))
        for v, _ in recs:
            table[v >> suf_bits] += 1
        for cnt in table:
            lut.append(n_rec)
            n_rec += cnt
        for v, c in recs:
            suf += (v & ((1 << suf_bits) - 1)).to_bytes(suf_bytes, "big")
            suf += (min(c, cmask)).to_bytes(counter_size, "little")
    lut.append(n_rec)                                   # guard entry
    suf += b"KMCS"
    pre = bytearray(b"KMCP")
    pre += struct.pack("<%dQ" % len(lut), *lut)
    if version == 0x200:
        nsig = (1 << (2 * signature_len)) + 1
        pre += struct.pack("<%dI" % nsig, *[(i * 7) % len(bins) for i in range(nsig)])
    hdr = struct.pack("<IIII", k, 0, counter_size, p)
    if version == 0x200:
        hdr += struct.pack("<I", signature_len)
    hdr += struct.pack("<IIQ", min_count, max_count & 0xFFFFFFFF, n_rec)
    hdr += struct.pack("<B3xI", 0 if canonical else 1, max_count >> 32)
    hdr += b"\0" * 20
    hdr += struct.pack("<I", version)
    pre += hdr
    pre += struct.pack("<I", len(hdr))
    pre += b"KMCP"
    with open(prefix + ".kmc_pre", "wb") as f:
        f.write(pre)
    with open(prefix + ".kmc_suf", "wb") as f:
        f.write(suf)


def read(prefix: str):
    """-> (dict(k=..., counter_size=..., min_count=..., max_count=..., total=..., canonical=..., version=...),
           list of (kmer string, count) in file order)"""
    with open(prefix + ".kmc_pre", "rb") as f:
        pre = f.read()
    with open(prefix + ".kmc_suf", "rb") as f:
        suf = f.read()
    if pre[:4] != b"KMCP" or pre[-4:] != b"KMCP" or suf[:4] != b"KMCS" or suf[-4:] != b"KMCS":
        raise ValueError("%s: KMC markers missing" % prefix)
    version, hoff = struct.unpack_from("<II", pre, len(pre) - 12)
    if version not in (0, 0x200):
        raise ValueError("%s: unknown kmc_version %#x" % (prefix, version))
    h0 = len(pre) - 8 - hoff
    k, mode, csz, p = struct.unpack_from("<IIII", pre, h0)
    o = h0 + 16
    sig = 0
    if version == 0x200:
        (sig,) = struct.unpack_from("<I", pre, o)
        o += 4
    mn, mx, total = struct.unpack_from("<IIQ", pre, o)
    o += 16
    nb, mxhi = struct.unpack_from("<B3xI", pre, o)
    lut_end = h0 - ((((1 << (2 * sig)) + 1) * 4) if version == 0x200 else 0)
    n_lut = (lut_end - 4) // 8
    lut = struct.unpack_from("<%dQ" % n_lut, pre, 4)
    suf_bytes = (k - p) // 4
    rec = suf_bytes + csz
    if 8 + total * rec != len(suf):
        raise ValueError("%s: .kmc_suf size does not match total_kmers" % prefix)
    out = []
    per = 1 << (2 * p)
    for i in range(n_lut - 1):
        pv = i % per
        for r in range(lut[i], lut[i + 1]):
            b = 4 + r * rec
            sv = int.from_bytes(suf[b:b + suf_bytes], "big")
            c = int.from_bytes(suf[b + suf_bytes:b + rec], "little")
            out.append((_to_str((pv << (2 * (k - p))) | sv, k), c))
    if len(out) != total:
        raise ValueError("%s: prefix tables do not cover total_kmers" % prefix)
    return dict(k=k, mode=mode, counter_size=csz, lut_prefix_length=p, min_count=mn, max_count=mx | (mxhi << 32),
                total=total, canonical=(nb == 0), version=version, signature_len=sig), out
