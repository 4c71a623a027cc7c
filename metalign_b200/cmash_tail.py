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
