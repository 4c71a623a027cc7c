"""The native read ingest (csrc/ingest.cpp) against the numpy restatement of the same record rules
(metalign_b200/ingest.py) and the host codec: FASTQ / FASTA, gzip, CRLF, no trailing newline, empty and long
reads, lower case, N runs, blocks and batches that cut records anywhere.  No GPU."""
import gzip
import os
import random

import numpy as np
import pytest

from metalign_b200 import codec, ingest


def _write(path, text: str):
    if str(path).endswith(".gz"):
        with gzip.open(path, "wt", newline="") as f:
            f.write(text)
    else:
        with open(path, "w", newline="") as f:
            f.write(text)


def _write_bgzf(path, data: bytes, rng, max_block=60000):
    """BGZF as bgzip / htslib write it: gzip members of at most 64 KiB with a 'BC' extra field holding the member size - 1,
    and the empty end-of-file member"""
    import struct
    import zlib

    def member(chunk):
        co = zlib.compressobj(6, zlib.DEFLATED, -15)
        body = co.compress(chunk) + co.flush()
        bsize = 12 + 6 + len(body) + 8 - 1
        return (b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, bsize) + body
                + struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk)))

    with open(path, "wb") as f:
        i = 0
        while i < len(data):
            n = rng.randint(1, max_block)
            f.write(member(data[i:i + n]))
            i += n
        f.write(member(b""))


def test_bgzf_members_inflate_in_parallel(tmp_path, monkeypatch):
    """a bgzip-style file goes through the parallel member path (and, with MLGI_NO_BGZF, through plain gzread): same reads"""
    rng = random.Random(77)
    reads = _random_reads(rng, 3000)
    text = "".join("@r%d\n%s\n+\n%s\n" % (i, r, "I" * len(r)) for i, r in enumerate(reads))
    p = tmp_path / "reads.fastq.gz"
    _write_bgzf(p, text.encode(), rng, max_block=5000)
    want = [_norm(r) for r in reads]
    assert gzip.open(p, "rb").read() == text.encode()            # a valid multi-member gzip file for everybody else
    for env, block, thr in (({}, 1 << 23, 0), ({}, 3000, 5), ({"MLGI_NO_BGZF": "1"}, 4096, 3)):
        monkeypatch.delenv("MLGI_NO_BGZF", raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        monkeypatch.setenv("MLGI_BLOCK_BYTES", str(block))
        got, _ = _native_reads(p, "fastq", reads_per_batch=700, threads=thr, bases_per_batch=200000)
        assert got == want, (env, block, thr)
    # a damaged member is an error, not silence
    raw = bytearray(p.read_bytes())
    raw[len(raw) // 2] ^= 0xFF
    bad = tmp_path / "bad.fastq.gz"
    bad.write_bytes(bytes(raw))
    monkeypatch.delenv("MLGI_NO_BGZF", raising=False)
    with pytest.raises(IOError):
        _native_reads(bad, "fastq", reads_per_batch=700, threads=4, bases_per_batch=200000)


def _native_reads(path, kind, **kw):
    rd = ingest.PackedBatches(str(path), kind, **kw)
    out = []
    nb = 0
    for bases, runs, off, n in rd:
        assert off[0] == 0 and len(off) == n + 1
        total = int(off[-1])
        mask = codec.runs_to_nmask(runs, total) if len(runs) else None
        # runs are sorted, disjoint and inside the batch
        if len(runs):
            ends = runs[:, 0].astype(np.int64) + runs[:, 1]
            assert (runs[:, 1] > 0).all() and ends[-1] <= total and (runs[1:, 0] > ends[:-1]).all()
        # padding after the last base is zero up to a 16-byte multiple
        used = (total + 3) // 4
        assert not bases[used:((used + 15) // 16) * 16 + 16].any()
        out += codec.unpack_reads(bases, mask, off)
        nb += 1
    st = rd.stats()
    rd.close()
    assert st["reads"] == len(out)
    return out, nb


def _numpy_reads(path, kind):
    out = []
    for text, off in ingest.batches(str(path), kind):
        s = text.tobytes().decode()
        out += [s[int(off[i]):int(off[i + 1])] for i in range(len(off) - 1)]
    return out


def _norm(r):
    return "".join(c if c in "ACGT" else "N" for c in r.upper())


def _random_reads(rng, n):
    reads = []
    for _ in range(n):
        L = rng.choice([0, 1, 3, 4, 5, 59, 60, 61, 150, 151, 250, 1000])
        r = [rng.choice("ACGT") for _ in range(L)]
        for i in range(L):
            x = rng.random()
            if x < 0.02:
                r[i] = rng.choice("NnRYKM.-*")
            elif x < 0.06:
                r[i] = r[i].lower()
        if L and rng.random() < 0.1:
            a = rng.randrange(L)
            for i in range(a, min(L, a + rng.randint(1, 80))):
                r[i] = "N"
        reads.append("".join(r))
    return reads


@pytest.mark.parametrize("ext,eol,trailing", [("fq", "\n", True), ("fastq.gz", "\n", False), ("fq", "\r\n", True)])
def test_fastq(tmp_path, monkeypatch, ext, eol, trailing):
    rng = random.Random(hash((ext, eol)) & 0xFFFF)
    reads = _random_reads(rng, 700)
    text = eol.join("@r%d desc\n%s\n+\n%s".replace("\n", eol) % (i, r, "I" * len(r)) for i, r in enumerate(reads))
    if trailing:
        text += eol
    p = tmp_path / ("reads." + ext)
    _write(p, text)
    want = [_norm(r) for r in reads]
    assert [_norm(r) for r in _numpy_reads(p, "fastq")] == want
    # plain files are mapped (views of the page cache), gzip ones are inflated into recycled blocks; MLGI_NO_MMAP /
    # MLGI_NO_AVX2 force the read(2) path and the SWAR packer / memchr scanner
    for block, rpb, thr, env in ((1 << 23, 4_000_000, 0, {}), (97, 50, 3, {}), (4096, 7, 1, {"MLGI_NO_MMAP": "1"}),
                                 (1000, 1000, 16, {"MLGI_NO_AVX2": "1"}), (97, 50, 3, {"MLGI_NO_MMAP": "1", "MLGI_NO_AVX2": "1"})):
        monkeypatch.setenv("MLGI_BLOCK_BYTES", str(block))
        for k in ("MLGI_NO_MMAP", "MLGI_NO_AVX2"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        got, nb = _native_reads(p, "fastq", reads_per_batch=rpb, threads=thr, bases_per_batch=max(2000, rpb * 200))
        assert got == want, (block, rpb, thr, env)
        assert nb >= len(reads) // rpb


def test_fasta_and_gz(tmp_path, monkeypatch):
    rng = random.Random(9)
    reads = [r for r in _random_reads(rng, 400) if r]            # an empty FASTA line is not a record
    lines = []
    for i, r in enumerate(reads):
        lines.append(">seq%d" % i)
        if i % 17 == 0:
            lines.append(";comment")
        if i % 23 == 0:
            lines.append("")
        lines.append(r)
    for name in ("reads.fa", "reads.fna.gz"):
        p = tmp_path / name
        _write(p, "\n".join(lines) + ("\n" if name.endswith("fa") else ""))
        want = [_norm(r) for r in reads]
        assert [_norm(r) for r in _numpy_reads(p, "fasta")] == want
        monkeypatch.setenv("MLGI_BLOCK_BYTES", "333")
        got, _ = _native_reads(p, "fasta", reads_per_batch=33, threads=4, bases_per_batch=40000)
        assert got == want


def test_limits_and_errors(tmp_path):
    p = tmp_path / "a.fq"
    _write(p, "@x\n" + "ACGT" * 100 + "\n+\n" + "I" * 400 + "\n")
    rd = ingest.PackedBatches(str(p), "fastq", reads_per_batch=10, bases_per_batch=100)
    with pytest.raises(IOError):
        next(rd)                                   # one read longer than a whole batch
    rd.close()
    with pytest.raises(IOError):
        ingest.PackedBatches(str(tmp_path / "missing.fq"), "fastq")
    e = tmp_path / "empty.fq"
    _write(e, "")
    assert _native_reads(e, "fastq") == ([], 0)
    # batches are cut by bases as well as by reads
    q = tmp_path / "b.fq"
    _write(q, "".join("@r\n%s\n+\n%s\n" % ("ACGTN" * 20, "I" * 100) for _ in range(50)))
    got, nb = _native_reads(q, "fastq", reads_per_batch=1000, bases_per_batch=1000)
    assert len(got) == 50 and nb == 5


def test_more_n_runs_than_the_buffer_holds(tmp_path):
    """low-quality reads: more separate N runs in a batch than the caller's run buffer -- KMC in the reference just skips
    N (scripts/select_db.py:50), so the batch must still come through (the reader keeps the runs and hands them over)"""
    p = tmp_path / "n.fq"
    read = "ACGNTTNACNGG" * 12                       # 36 N runs per 144-base read
    _write(p, "".join("@r%d\n%s\n+\n%s\n" % (i, read, "I" * len(read)) for i in range(400)))
    rd = ingest.PackedBatches(str(p), "fastq", reads_per_batch=1000, bases_per_batch=100000, max_runs=64)
    bases, runs, off, n = next(rd)
    assert n == 400 and runs.shape == (400 * 36, 2) and (runs[:, 1] == 1).all()
    want = np.array([i for i, c in enumerate(read * 400) if c == "N"], dtype=np.uint32)
    assert np.array_equal(runs[:, 0], want)
    with pytest.raises(StopIteration):
        next(rd)
    rd.close()


def test_large_parallel_pack_matches_codec(tmp_path):
    """~6 Mbases through 8 workers: the packed stream and the run list equal codec.pack_reads exactly"""
    rng = np.random.default_rng(5)
    n, L = 40000, 150
    arr = rng.integers(0, 4, size=(n, L), dtype=np.uint8)
    chars = np.frombuffer(b"ACGT", dtype=np.uint8)[arr]
    chars[rng.random((n, L)) < 0.001] = ord("N")
    reads = [bytes(row).decode() for row in chars]
    p = tmp_path / "big.fq"
    _write(p, "".join("@r\n%s\n+\n%s\n" % (r, "I" * L) for r in reads))
    rd = ingest.PackedBatches(str(p), "fastq", reads_per_batch=n + 5, threads=8)
    bases, runs, off, m = next(rd)
    assert m == n
    wb, wm, woff = codec.pack_reads(reads)
    total = n * L
    assert np.array_equal(off, woff)
    assert np.array_equal(bases[: (total + 3) // 4], wb[: (total + 3) // 4])
    assert np.array_equal(runs, codec.nmask_to_runs(wm, total))
    with pytest.raises(StopIteration):
        next(rd)
    rd.close()


def test_fast_inflate_against_zlib():
    """csrc/fast_inflate.h (the gzip decoder of the ingest) against zlib: every compression level (0 = stored blocks, tiny
    inputs = fixed Huffman), random / repetitive / text-like data, several members in one file, header fields, and stopping
    and resuming at every few output bytes; damaged streams fail (or fail their CRC) instead of crashing"""
    import ctypes as C
    import zlib
    L = ingest.ingest_lib()
    L.mlgi_test_inflate.argtypes = [C.c_char_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64)]
    rng = random.Random(2024)

    def gz(data, level, name=b""):
        co = zlib.compressobj(level, zlib.DEFLATED, -15)
        body = co.compress(data) + co.flush()
        flg = 8 if name else 0
        hdr = b"\x1f\x8b\x08" + bytes([flg]) + b"\0\0\0\0\0\xff" + (name + b"\0" if name else b"")
        return hdr + body + (zlib.crc32(data) & 0xFFFFFFFF).to_bytes(4, "little") + (len(data) & 0xFFFFFFFF).to_bytes(4, "little")

    def inflate(blob, want_len, step):
        out = np.empty(want_len + step + 1024, dtype=np.uint8)
        n = C.c_uint64()
        rc = L.mlgi_test_inflate(blob, len(blob), out.ctypes.data, out.size, step, C.byref(n))
        return rc, out[:n.value].tobytes() if rc == 0 else b""

    def sample(kind, n):
        if kind == "random":
            return bytes(rng.randrange(256) for _ in range(n))
        if kind == "dna":
            return "".join("@r%d\n%s\n+\n%s\n" % (i, "".join(rng.choice("ACGT") for _ in range(100)),
                                                   "".join(rng.choice("FFFF:,#") for _ in range(100))) for i in range(n // 210 + 1)).encode()[:n]
        if kind == "runs":
            return b"".join(bytes([rng.randrange(4) + 65]) * rng.randint(1, 700) for _ in range(n // 300 + 1))[:n]
        return (b"the quick brown fox jumps over the lazy dog " * (n // 44 + 1))[:n]

    cases = 0
    for kind in ("random", "dna", "runs", "text"):
        for n in (0, 1, 5, 300, 70000, 400000):
            data = sample(kind, n)
            for level in (0, 1, 6, 9):
                blob = gz(data, level, name=b"x.fq" if level == 6 else b"")
                for step in (1 << 22, 7, 1000):
                    if step == 7 and n > 70000:
                        continue
                    rc, got = inflate(blob, len(data), step)
                    assert rc == 0 and got == data, (kind, n, level, step, rc)
                    cases += 1
    # several members, and zero padding behind the last one
    parts = [sample("dna", 5000), b"", sample("text", 100000), sample("random", 3000)]
    blob = b"".join(gz(p, lv) for p, lv in zip(parts, (1, 6, 0, 9))) + b"\0" * 37
    rc, got = inflate(blob, sum(map(len, parts)), 4096)
    assert rc == 0 and got == b"".join(parts)
    # damage: never a crash, never a silent success with wrong bytes
    data = sample("dna", 200000)
    blob = bytearray(gz(data, 6))
    silent = 0
    for _ in range(300):
        b2 = bytearray(blob)
        for _ in range(rng.randint(1, 3)):
            b2[rng.randrange(len(b2))] ^= 1 << rng.randrange(8)
        rc, got = inflate(bytes(b2), len(data) + 70000, 1 << 16)
        if rc == 0:
            assert got == data             # only flips in ignored header fields (mtime, xfl, os) can leave the data intact
            silent += 1
    assert silent < 30
    rc, _ = inflate(bytes(blob[:len(blob) // 2]), len(data), 1 << 16)
    assert rc != 0
    assert cases > 200


def test_fast_inflate_truncated_at_every_byte():
    """A gzip stream cut at ANY byte must fail cleanly -- never run past the output cap, never hang.  The data are chosen
    for long codes: Z_HUFFMAN_ONLY over a skewed alphabet gives literal codes beyond the 11-bit primary table, and a
    level-9 stream over far-apart repeats gives distance codes beyond the 8-bit one (the second-level paths, where the
    unsigned bit counter once wrapped on a truncated tail)."""
    import ctypes as C
    import zlib
    L = ingest.ingest_lib()
    L.mlgi_test_inflate.argtypes = [C.c_char_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64)]
    rng = random.Random(77)

    def gz(data, level, strategy):
        co = zlib.compressobj(level, zlib.DEFLATED, -15, 9, strategy)
        body = co.compress(data) + co.flush()
        return (b"\x1f\x8b\x08\0\0\0\0\0\0\xff" + body + (zlib.crc32(data) & 0xFFFFFFFF).to_bytes(4, "little")
                + (len(data) & 0xFFFFFFFF).to_bytes(4, "little"))

    # skewed alphabet: byte value b with probability ~ 2^-(b/8)
    skew = bytes(min(255, int(rng.expovariate(0.09))) for _ in range(40000))
    chunks = [bytes(rng.randrange(256) for _ in range(rng.randint(20, 200))) for _ in range(300)]
    far = b"".join(rng.choice(chunks) for _ in range(1500))[:90000]
    total = 0
    for data, level, strategy in ((skew, 6, zlib.Z_HUFFMAN_ONLY), (far, 9, zlib.Z_DEFAULT_STRATEGY)):
        blob = gz(data, level, strategy)
        out = np.empty(len(data) + 3 * 4096, dtype=np.uint8)
        n = C.c_uint64()
        assert L.mlgi_test_inflate(blob, len(blob), out.ctypes.data, out.size, 4096, C.byref(n)) == 0
        assert out[:n.value].tobytes() == data
        for cut in range(1, len(blob)):          # (an empty file is a valid gzip stream of zero members)
            rc = L.mlgi_test_inflate(blob[:cut], cut, out.ctypes.data, out.size, 4096, C.byref(n))
            assert rc in (-1, -3), (level, cut, rc)
            total += 1
    assert total > 50000
