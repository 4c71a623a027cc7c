#!/usr/bin/env python
"""Generates tests/golden/hot_path_cases.json: hand-derived known-answer cases for the hot path.

The reference tree holds no golden vector for this path and KMC / CMash are not
installed, so these cases are the repo's own pins.  The EXPECTED tables below are written by hand from
the semantics in SURVEY.md 3.3 (the derivation is in each case's `why`); nothing here calls the oracle or
the CUDA path.  Sequences are pseudo-random (seeded) so that no accidental overlap exists between them.

Run:  python tests/golden/make_hot_path_cases.py
"""
import json
import os
import random

K = 60
KS = [30, 40, 50, 60]
_COMP = str.maketrans("ACGT", "TGCA")


def rc(s):
    return s.translate(_COMP)[::-1]


rng = random.Random(20200529)


def rnd(n):
    return "".join(rng.choice("ACGT") for _ in range(n))


def fwd_canonical(n=K):
    """random n-mer that is its own canonical form (s < rc(s))"""
    while True:
        s = rnd(n)
        if s < rc(s):
            return s


def rc_canonical(n=K):
    """random n-mer whose canonical form is its reverse complement"""
    s = fwd_canonical(n)
    return rc(s)


def embed(kmer, left=23, right=31):
    return rnd(left) + kmer + rnd(right)


cases = []


def case(name, why, sketches, reads, expect):
    cases.append(dict(name=name, why=why, K=K, ks=KS, sketches=sketches, reads=reads, expect=expect))


def exp(num, den, I, gate="exact", count_empty=True, ci_min=2):
    return dict(gate=gate, count_empty_in_den=count_empty, ci_min=ci_min, num=num, den=den, I=sorted(I))


# 1 ---------------------------------------------------------------------------------------------------
s1, s2 = fwd_canonical(), fwd_canonical()
case("forward_canonical_present",
     "s1 is its own canonical form and occurs in two reads: I={s1}; offset 0 forward-matches slot 0 at k=30, "
     "the gate passes, and x[0:40], x[0:50], x[0:60] forward-match too -> 1,1,1,1.  s2 never occurs.",
     [[s1, s2]], [embed(s1), embed(s1, 5, 40)],
     [exp([[1, 1, 1, 1]], [[2, 2, 2, 2]], [s1], g) for g in ("exact", "none")])

# 2 ---------------------------------------------------------------------------------------------------
r = rc_canonical()
case("rc_canonical_present",
     "the sketch stores r but KMC emits x=rc(r).  Only the reverse-complement lookup of x[30:60] equals r[:30] "
     "(k=30 hit at offset 30, where no longer k fits).  k=40/50/60 would match at offsets 20/10/0 but the "
     "30-mers there are not in the prefilter -> 1,0,0,0 gated, 1,1,1,1 ungated.",
     [[r]], [embed(r), embed(r, 40, 7)],
     [exp([[1, 0, 0, 0]], [[1, 1, 1, 1]], [rc(r)], "exact"), exp([[1, 1, 1, 1]], [[1, 1, 1, 1]], [rc(r)], "none")])

# 3 ---------------------------------------------------------------------------------------------------
s = fwd_canonical()
case("singleton_absent", "s occurs once in the whole read set: dropped by -ci2.",
     [[s]], [embed(s), rnd(150)],
     [exp([[0, 0, 0, 0]], [[1, 1, 1, 1]], [], g) for g in ("exact", "none")])
case("singleton_present_with_ci1", "same input, ci_min=1 keeps it.",
     [[s]], [embed(s), rnd(150)],
     [exp([[1, 1, 1, 1]], [[1, 1, 1, 1]], [s], "exact", ci_min=1)])
case("both_strands_pool", "s once forward and once reverse-complemented: canonical counting pools them -> count 2.",
     [[s]], [embed(s), embed(rc(s), 11, 17)],
     [exp([[1, 1, 1, 1]], [[1, 1, 1, 1]], [s], "exact")])
case("tandem_repeat_in_one_read", "s twice in ONE read: occurrences are counted, not reads.",
     [[s]], [rnd(9) + s + s + rnd(9)],
     [exp([[1, 1, 1, 1]], [[1, 1, 1, 1]], [s], "exact")])

# 4 ---------------------------------------------------------------------------------------------------
s, a, b = fwd_canonical(), fwd_canonical(), fwd_canonical()
case("shared_kmer_same_orientation", "two strains store the same 60-mer: one D entry, both credited.",
     [[s, a], [s, b]], [embed(s), embed(s, 3, 3)],
     [exp([[1, 1, 1, 1], [1, 1, 1, 1]], [[2, 2, 2, 2], [2, 2, 2, 2]], [s], g) for g in ("exact", "none")])
case("shared_kmer_opposite_orientation",
     "genome 1 stores rc(s).  x=s.  Offset 0 forward-matches genome 0 at every k, and forward-first means the "
     "reverse-complement lookup is never tried there, so genome 1 never gets k=60.  Genome 1 gets k=30 at offset 30 "
     "(reverse complement, forward empty).  Ungated it also gets k=40 at offset 20 and k=50 at offset 10 "
     "(forward empty there); gated those offsets fail the prefilter.",
     [[s, a], [rc(s), b]], [embed(s), embed(s, 3, 3)],
     [exp([[1, 1, 1, 1], [1, 0, 0, 0]], [[2, 2, 2, 2], [2, 2, 2, 2]], [s], "exact"),
      exp([[1, 1, 1, 1], [1, 1, 1, 0]], [[2, 2, 2, 2], [2, 2, 2, 2]], [s], "none")])

# 5 ---------------------------------------------------------------------------------------------------
s = fwd_canonical()
t = s[10:40] + rnd(30)
case("cross_genome_30_prefix",
     "genome 1's k-mer t starts with s[10:40]; t itself never occurs in the reads.  The window of x=s at offset 10 "
     "forward-matches t at k=30 only -> genome 1 gets 1,0,0,0 (and would be dropped by the k=60 > 0 filter).",
     [[s], [t]], [embed(s), embed(s, 2, 9)],
     [exp([[1, 1, 1, 1], [1, 0, 0, 0]], [[1, 1, 1, 1], [1, 1, 1, 1]], [s], g) for g in ("exact", "none")])

# 6 ---------------------------------------------------------------------------------------------------
s = fwd_canonical()
last = s[50:]
alt = last
while alt == last:
    alt = rnd(10)
s_alt = s[:50] + alt
case("duplicate_prefixes_in_one_sketch",
     "slots 0 and 1 share their first 50 bases: both are hit at k=30/40/50 but count as ONE prefix; the "
     "denominator dedupes the same way -> num 1,1,1,1 over den 1,1,1,2.",
     [[s, s_alt]], [embed(s), embed(s, 8, 8)],
     [exp([[1, 1, 1, 1]], [[1, 1, 1, 2]], [s], g) for g in ("exact", "none")])

# 7 ---------------------------------------------------------------------------------------------------
s = fwd_canonical()
case("padded_sketch_denominator",
     "a sketch with unused ('') slots: CMash's len({kmer[:k] ...}) counts '' once.",
     [[s, "", ""]], [embed(s), embed(s, 8, 8)],
     [exp([[1, 1, 1, 1]], [[2, 2, 2, 2]], [s], "exact", True), exp([[1, 1, 1, 1]], [[1, 1, 1, 1]], [s], "exact", False)])

# 8 ---------------------------------------------------------------------------------------------------
s = fwd_canonical()
withN = s[:17] + "N" + s[18:]
case("n_lowercase_short",
     "lower case is folded (counts), a read with an N inside the k-mer and a 59-base read do not count: "
     "lower + upper = 2 -> present.",
     [[s]], [embed(s.lower()), embed(withN), s[:59], embed(s, 1, 1)],
     [exp([[1, 1, 1, 1]], [[1, 1, 1, 1]], [s], "exact")])
case("n_breaks_window",
     "only the lower-case copy is intact; the N copy and the truncated copy do not count -> count 1 -> absent.",
     [[s]], [embed(s.lower()), embed(withN), s[:59]],
     [exp([[0, 0, 0, 0]], [[1, 1, 1, 1]], [], "exact")])

# 9 ---------------------------------------------------------------------------------------------------
h = rnd(30)
pal = h + rc(h)
assert pal == rc(pal)
case("palindrome", "x == rc(x): one canonical key; two occurrences -> present; offset 0 forward-matches at every k.",
     [[pal]], [embed(pal), embed(pal, 4, 4)],
     [exp([[1, 1, 1, 1]], [[1, 1, 1, 1]], [pal], g) for g in ("exact", "none")])
case("palindrome_single", "a palindromic k-mer seen once counts ONCE (not once per strand).",
     [[pal]], [embed(pal)],
     [exp([[0, 0, 0, 0]], [[1, 1, 1, 1]], [], "exact")])

out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hot_path_cases.json")
with open(out, "w") as f:
    json.dump(cases, f, indent=1)
print("wrote %d cases to %s" % (len(cases), out))
