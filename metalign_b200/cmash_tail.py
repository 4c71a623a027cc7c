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
