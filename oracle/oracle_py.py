"""TEST INFRASTRUCTURE -- CPU oracle, set-based Python restatement.  PARITY UNPINNED.

This file restates, on plain Python strings / dicts / sets, the semantics of the
database-selection hot path of nlapier2/Metalign:

    scripts/select_db.py:43-65   run_kmc_steps      (kmc -k60 -ci2 -cs3, kmc_tools intersect, kmc_dump)
    scripts/select_db.py:68-76   run_cmash_and_cutoff (StreamingQueryDNADatabase.py ... 30-60-10 -c 0 --sensitive)
    scripts/select_db.py:80-96   cutoff + one-strain-per-species selection
    local_tests/dump_kmers.py:7-14 and local_tests/retrain_and_test_metalign.sh:49-66 (DB-side k-mer set D)

The arithmetic of that path lives in two third-party programs that are NOT in the
reference tree and are not installed here: KMC 3.x (Refresh-Bioinformatics) and
CMash (dkoslicki/CMash, StreamingQueryDNADatabase.py / MinHash.py).  The reference
pins no version of either (setup.py:13-37 has no install_requires) and holds no
golden vector or test for this path.  The oracle therefore follows the published
behaviour of those tools as written down in SURVEY.md section 3.3 (steps R1-R7),
and exposes every uncertain behaviour as a switch:

    gate               'exact' (zero-false-positive Bloom prefilter) | 'none'
    count_empty_in_den True: a sketch with unused ('') slots counts '' once in the denominator
    ci_min             2  (kmc -ci2)

"PARITY UNPINNED": no output of the real KMC/CMash binaries was available to check
this file against; it is pinned only by the hand-derived cases in tests/golden/.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this
module.  It is the checker, never the product path.
"""
from __future__ import annotations

from collections import Counter, defaultdict
from typing import Dict, Iterable, List, Sequence, Set, Tuple

_COMP = str.maketrans("ACGT", "TGCA")
_VALID = frozenset("ACGT")


def rc(s: str) -> str:
    """Reverse complement of an upper-case ACGT string."""
    return s.translate(_COMP)[::-1]


def canon(s: str) -> str:
    """KMC canonical form: lexicographic min of k-mer and its reverse complement
    (A<C<G<T; equal to the numeric min of the 2-bit encodings).  SURVEY.md A.1."""
    r = rc(s)
    return s if s <= r else r


# --------------------------------------------------------------------------- R1
def count_read_kmers(reads: Iterable[str], K: int = 60) -> Counter:
    """R1 -- `kmc -k60 -ci2 -cs3` (select_db.py:50-52), counting part.

    Every K-long window of every maximal ACGT run of every read is counted under
    its canonical form; lower case is folded; any other symbol breaks the run.
    """
    cnt: Counter = Counter()
    for r in reads:
        r = r.upper()
        piece_start = 0
        n = len(r)
        i = 0
        while i <= n:
            if i == n or r[i] not in _VALID:
                piece = r[piece_start:i]
                for p in range(len(piece) - K + 1):
                    cnt[canon(piece[p:p + K])] += 1
                piece_start = i + 1
            i += 1
    return cnt


# --------------------------------------------------------------------------- R2
def db_kmer_set(sketches: Sequence[Sequence[str]]) -> Set[str]:
    """R2 -- D = canonical form of every non-empty sketch slot
    (dump_kmers.py:10-14 -> `kmc -k60 -fa -ci0`, retrain_and_test_metalign.sh:66)."""
    return {canon(s) for sk in sketches for s in sk if s != ""}


# --------------------------------------------------------------------------- R3
def intersect(cnt: Counter, D: Set[str], ci_min: int = 2) -> Set[str]:
    """R3 -- I = {x : cnt[x] >= ci_min} & D  (select_db.py:54-65)."""
    return {x for x in D if cnt.get(x, 0) >= ci_min}


# --------------------------------------------------------------------------- R4
class _Trie:
    """Stand-in for marisa-trie `keys(prefix)` over the sketch k-mers: a dict from
    every ks-length prefix to the (g, j) slots that start with it."""

    def __init__(self, sketches: Sequence[Sequence[str]], ks: Sequence[int]):
        self.by_prefix: Dict[str, List[Tuple[int, int]]] = defaultdict(list)
        for g, sk in enumerate(sketches):
            for j, s in enumerate(sk):
                if s == "":
                    continue
                for k in ks:
                    if k <= len(s):
                        self.by_prefix[s[:k]].append((g, j))

    def prefix_matches(self, c: str) -> List[Tuple[int, int]]:
        return self.by_prefix.get(c, [])

    def match(self, w: str) -> List[Tuple[int, int]]:
        """forward first; reverse complement ONLY if forward found nothing."""
        m = self.prefix_matches(w)
        if m:
            return m
        return self.prefix_matches(rc(w))


def query_hits(I: Iterable[str], sketches: Sequence[Sequence[str]], ks: Sequence[int],
               gate: str = "exact") -> Set[Tuple[int, int, int]]:
    """R4 -- the per-record loop of StreamingQueryDNADatabase.py on the records of I.
    Returns H = {(g, k, j)}."""
    assert gate in ("exact", "none")
    ks = list(ks)
    k0 = ks[0]
    trie = _Trie(sketches, ks)
    E0: Set[str] = set()
    if gate == "exact":
        for sk in sketches:
            for s in sk:
                if s != "":
                    E0.add(s[:k0])
                    E0.add(rc(s[:k0]))
    H: Set[Tuple[int, int, int]] = set()
    for x in I:
        L = len(x)
        for i in range(L - k0 + 1):
            w0 = x[i:i + k0]
            possible = True if gate == "none" else (w0 in E0)
            if not possible:
                continue
            for (g, j) in trie.match(w0):
                H.add((g, k0, j))
            for k in ks[1:]:
                if i + k > L:
                    continue
                wk = x[i:i + k]
                for (g, j) in trie.match(wk):
                    H.add((g, k, j))
    return H


# --------------------------------------------------------------------------- R5
def containment_table(H: Set[Tuple[int, int, int]], sketches: Sequence[Sequence[str]],
                      ks: Sequence[int], count_empty_in_den: bool = True):
    """R5 -- distinct hit prefixes / distinct sketch prefixes, per genome and k.
    Returns (num, den, ci) as lists of G rows of len(ks) entries."""
    G = len(sketches)
    ks = list(ks)
    hit_prefixes: Dict[Tuple[int, int], Set[str]] = defaultdict(set)
    for (g, k, j) in H:
        hit_prefixes[(g, k)].add(sketches[g][j][:k])
    num = [[0] * len(ks) for _ in range(G)]
    den = [[0] * len(ks) for _ in range(G)]
    ci = [[0.0] * len(ks) for _ in range(G)]
    for g in range(G):
        for ki, k in enumerate(ks):
            if count_empty_in_den:
                d = len({s[:k] for s in sketches[g]})
            else:
                d = len({s[:k] for s in sketches[g] if s != ""})
            den[g][ki] = d
            nn = len(hit_prefixes.get((g, k), ()))
            num[g][ki] = nn
            ci[g][ki] = (float(nn) / float(d)) if nn > 0 else 0.0
    return num, den, ci


# --------------------------------------------------------------------------- R6', only without --sensitive
def refilter_unique(H: Set[Tuple[int, int, int]], sketches: Sequence[Sequence[str]], ks: Sequence[int], ci,
                    coverage_threshold: float = 0.0):
    """CMash's post-processing when --sensitive is ABSENT [UPSTREAM, SURVEY.md A.2: "re-filters to k-mers unique to one
    organism"; Metalign never takes this branch, select_db.py:76].  UNPINNED restatement, as read here:
      * the organisms that pass the basic filter (containment at the largest k > threshold) are the candidates;
      * at every k, a k-prefix of a candidate's sketch is UNIQUE when no other candidate's sketch has a k-mer with that
        prefix ('' slots take no part);
      * a candidate's containment is recomputed over its unique prefixes only: hit unique prefixes / unique prefixes
        (0.0 where it has none).
    Returns (candidates in genome order, num, den, ci) with one row per candidate."""
    ks = list(ks)
    cand = [g for g in range(len(sketches)) if ci[g][len(ks) - 1] > coverage_threshold]
    num, den, out = [], [], []
    for ki, k in enumerate(ks):
        owners: Dict[str, Set[int]] = defaultdict(set)
        for g in cand:
            for s in sketches[g]:
                if s != "":
                    owners[s[:k]].add(g)
        hit = defaultdict(set)
        for (g, kk, j) in H:
            if kk == k:
                hit[g].add(sketches[g][j][:k])
        for row, g in enumerate(cand):
            if ki == 0:
                num.append([0] * len(ks)); den.append([0] * len(ks)); out.append([0.0] * len(ks))
            uniq = {s[:k] for s in sketches[g] if s != "" and len(owners[s[:k]]) == 1}
            n_hit = len(uniq & hit[g])
            num[row][ki], den[row][ki] = n_hit, len(uniq)
            out[row][ki] = (float(n_hit) / float(len(uniq))) if uniq and n_hit else 0.0
    return cand, num, den, out


def run(reads: Iterable[str], sketches: Sequence[Sequence[str]], K: int = 60,
        ks: Sequence[int] = (30, 40, 50, 60), ci_min: int = 2, gate: str = "exact",
        count_empty_in_den: bool = True):
    """R1-R5 end to end.  Returns dict(I=sorted list, num, den, ci, n_kmers)."""
    cnt = count_read_kmers(reads, K)
    D = db_kmer_set(sketches)
    I = intersect(cnt, D, ci_min)
    H = query_hits(I, sketches, ks, gate)
    num, den, ci = containment_table(H, sketches, ks, count_empty_in_den)
    return dict(I=sorted(I), num=num, den=den, ci=ci, n_kmers=sum(cnt.values()))


# --------------------------------------------------------------------------- R7
def select_organisms(rows: Sequence[Tuple[str, float]], taxid2info: Dict[str, list],
                     cutoff: float = 0.01, strain_level: bool = False) -> List[str]:
    """R7 -- select_db.py:80-96 on (name, last-column containment) rows in CSV order."""
    out: List[str] = []
    species_included: Dict[str, int] = {}
    for organism, containment_index in rows:
        if containment_index >= cutoff:
            if not strain_level:
                taxid = organism.split("taxid_")[1].split("_genomic.fna")[0].replace("_", ".")
                species = taxid2info[taxid][3].split("|")[-2]
                if species not in species_included or species == "":
                    species_included[species] = 1
                else:
                    continue
            out.append(organism)
    return out
