"""Host codec round trips (no GPU): packed stream, N mask, N runs."""
import random

import numpy as np

from metalign_b200 import codec


def test_pack_unpack_round_trip():
    rng = random.Random(3)
    reads = ["".join(rng.choice("ACGTN") for _ in range(rng.randint(0, 90))) for _ in range(50)]
    bases, nmask, off = codec.pack_reads(reads)
    assert bases.size % 16 == 0 and nmask.size % 16 == 0
    assert codec.unpack_reads(bases, nmask, off) == reads


def test_nmask_runs_round_trip():
    rng = random.Random(4)
    for _ in range(20):
        n = rng.randint(1, 700)
        isn = np.zeros(n, dtype=np.uint8)
        for _ in range(rng.randint(0, 12)):
            a = rng.randint(0, n - 1)
            isn[a:a + rng.randint(1, 70)] = 1
        mask = np.packbits(isn)
        runs = codec.nmask_to_runs(mask, n)
        assert runs.dtype == np.uint32 and runs.shape[1] == 2
        # sorted, non-overlapping, non-adjacent, inside the stream
        ends = runs[:, 0].astype(np.int64) + runs[:, 1]
        assert (runs[:, 1] > 0).all() and (ends <= n).all()
        assert (runs[1:, 0].astype(np.int64) > ends[:-1]).all()
        assert int(runs[:, 1].sum()) == int(isn.sum())
        back = codec.runs_to_nmask(runs, n)
        assert np.array_equal(back[: mask.size], mask)


def test_empty_runs():
    runs = codec.nmask_to_runs(np.zeros(16, dtype=np.uint8), 100)
    assert runs.shape == (0, 2)
