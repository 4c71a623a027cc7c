"""The tail end of CMash's StreamingQueryDNADatabase.py as Metalign runs it (scripts/select_db.py:73-76 of the
reference: `... 30-60-10 -c 0 ... --sensitive`), i.e. SURVEY.md 3.3 step R6.  It stays on the host and uses the
same pandas calls, because the row ORDER of the CSV (an unstable sort of a column full of ties) decides
which strain of a species select_db keeps (scripts/select_db.py:80-96)."""
from __future__ import annotations

from typing import Sequence

import numpy as np
import pandas as pd


def containment_frame(names: Sequence[str], ks: Sequence[int], ci: np.ndarray) -> pd.DataFrame:
    """All genomes x all k, columns 'k=30' ... in the order CMash builds them, index = sketch names."""
    ci = np.asarray(ci, dtype=np.float64)
    data = {"k=%d" % k: ci[:, i] for i, k in enumerate(ks)}
    return pd.DataFrame(data, index=list(names))


def filter_and_sort(df: pd.DataFrame, coverage_threshold: float = 0.0) -> pd.DataFrame:
    """-c <threshold> applied with a strict '>' at the largest k (CMash's default location), then
    sort_values on that column, descending, default (unstable) sort kind."""
    max_key = df.columns[-1]
    kept = df[df[max_key] > coverage_threshold]
    return kept.sort_values(max_key, ascending=False)


def write_results_csv(path: str, names: Sequence[str], ks: Sequence[int], ci: np.ndarray,
                      coverage_threshold: float = 0.0) -> pd.DataFrame:
    out = filter_and_sort(containment_frame(names, ks, ci), coverage_threshold)
    out.to_csv(path, index=True, encoding="utf-8")
    return out


def write_results_csv_sparse(path: str, names: Sequence[str], ks: Sequence[int], genomes: np.ndarray, ci_rows: np.ndarray,
                             coverage_threshold: float = 0.0) -> pd.DataFrame:
    """The same file from the sparse form of the result (Query.finish_sparse: rows only for genomes with a hit, in genome
    order).  CMash filters BEFORE it sorts, so the sort sees exactly the rows above the threshold in genome order either
    way, and with a threshold >= 0 every such row has a hit: the two forms give the same bytes (tests/test_select_db.py)."""
    if coverage_threshold < 0.0:
        raise ValueError("the sparse form only holds genomes with a hit: it cannot serve a negative threshold")
    ci_rows = np.asarray(ci_rows, dtype=np.float64).reshape(-1, len(ks))
    data = {"k=%d" % k: ci_rows[:, i] for i, k in enumerate(ks)}
    df = pd.DataFrame(data, index=[names[int(g)] for g in genomes])
    out = filter_and_sort(df, coverage_threshold)
    out.to_csv(path, index=True, encoding="utf-8")
    return out


# ---- CMash's post-processing when --sensitive is ABSENT.  Metalign always passes --sensitive
# (scripts/select_db.py:76), so the drop-in never comes here; this is for callers that use the library as a CMash query
# engine.  CMash does this step on the host, in Python, over the handful of organisms that passed the basic filter; so
# does this.  UNPINNED restatement of [UPSTREAM] behaviour ("re-filters to k-mers unique to one organism"),
# checked against oracle/oracle_py.py: refilter_unique.
def _prefix_ids(keys: np.ndarray, K: int, k: int) -> np.ndarray:
    """the k leading bases of (hi, lo) K-mer keys as a structured array that np.unique can group by"""
    hi, lo = keys[:, 0].astype(np.uint64), keys[:, 1].astype(np.uint64)
    s = 2 * (K - k)
    if s == 0:
        phi, plo = hi, lo
    elif s < 64:
        plo = (lo >> np.uint64(s)) | (hi << np.uint64(64 - s))
        phi = hi >> np.uint64(s)
    else:
        plo = hi >> np.uint64(s - 64) if s > 64 else hi.copy()
        phi = np.zeros_like(hi)
    out = np.empty(keys.shape[0], dtype=[("hi", "<u8"), ("lo", "<u8")])
    out["hi"], out["lo"] = phi, plo
    return out


def refilter_unique(keys_sel: np.ndarray, hit_flags: np.ndarray, K: int, ks: Sequence[int]):
    """keys_sel: (m, n, 2) uint64 source keys of the m candidate genomes (EMPTY = (~0, ~0) for '' slots);
    hit_flags: (m, nk, n) uint8 from Query.hit_flags (1 at the representative slot of every hit class).
    Returns (num, den, ci), each (m, nk): hit unique prefixes, unique prefixes, their ratio (0.0 without hits)."""
    keys_sel = np.asarray(keys_sel, dtype=np.uint64)
    m, n = keys_sel.shape[0], keys_sel.shape[1]
    nk = len(ks)
    num = np.zeros((m, nk), dtype=np.int64)
    den = np.zeros((m, nk), dtype=np.int64)
    flat = keys_sel.reshape(m * n, 2)
    real = ~((flat[:, 0] == np.uint64(0xFFFFFFFFFFFFFFFF)) & (flat[:, 1] == np.uint64(0xFFFFFFFFFFFFFFFF)))
    genome = np.repeat(np.arange(m), n)
    for ki, k in enumerate(ks):
        pid = _prefix_ids(flat, K, int(k))
        _, cls = np.unique(pid[real], return_inverse=True)           # class id of every real slot, shared across genomes
        g_real = genome[real]
        hit_real = hit_flags[:, ki, :].reshape(m * n)[real].astype(bool)
        ncls = int(cls.max()) + 1 if cls.size else 0
        # (class, genome) pairs: how many candidate genomes own each class, and whether a genome's copy of it was hit
        pair = cls.astype(np.int64) * m + g_real
        upair, inv = np.unique(pair, return_inverse=True)
        owners = np.bincount(upair // m, minlength=ncls)             # genomes per class
        pair_hit = np.zeros(upair.size, dtype=bool)
        np.logical_or.at(pair_hit, inv, hit_real)                    # the flag sits on one slot of the class: OR over its slots
        unique_pair = owners[upair // m] == 1
        pg = (upair % m).astype(np.int64)
        den[:, ki] = np.bincount(pg[unique_pair], minlength=m)
        num[:, ki] = np.bincount(pg[unique_pair & pair_hit], minlength=m)
    with np.errstate(divide="ignore", invalid="ignore"):
        ci = np.where((num > 0) & (den > 0), num.astype(np.float64) / np.maximum(den, 1).astype(np.float64), 0.0)
    return num, den, ci


def write_results_csv_specific(path: str, names: Sequence[str], ks: Sequence[int], K: int, genomes: np.ndarray, ci_rows: np.ndarray,
                               keys_of, hit_flags_of, coverage_threshold: float = 0.0) -> pd.DataFrame:
    """The CSV CMash writes WITHOUT --sensitive: basic filter first (rows of `genomes` whose containment at the largest k
    exceeds the threshold), unique-prefix re-filter over those candidates, threshold and sort again on the re-computed column.
    keys_of(candidates) -> (m, n, 2) source keys (dbformat.read_keys_rows on the source-form .mlgdb);
    hit_flags_of(candidates) -> (m, nk, n) flags (Query.hit_flags)."""
    ci_rows = np.asarray(ci_rows, dtype=np.float64).reshape(-1, len(ks))
    genomes = np.asarray(genomes)
    keep = ci_rows[:, -1] > coverage_threshold
    cand = genomes[keep]
    order = np.argsort(cand, kind="stable")
    cand = cand[order]
    _, _, ci2 = refilter_unique(keys_of(cand), hit_flags_of(cand), K, ks) if cand.size else (None, None, np.zeros((0, len(ks))))
    data = {"k=%d" % k: ci2[:, i] for i, k in enumerate(ks)}
    df = pd.DataFrame(data, index=[names[int(g)] for g in cand])
    out = filter_and_sort(df, coverage_threshold)
    out.to_csv(path, index=True, encoding="utf-8")
    return out
