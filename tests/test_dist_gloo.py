"""The multi-rank path on CPU: world_size 2 over gloo.  The exchange step (shard the reads, clamp each rank's
per-k-mer counters, ONE uint8 sum all-reduce, derive presence from the sum) is the host logic of
metalign_b200.dist; the engine behind it here is the CPU oracle, which exposes the same export/import seam
as the GPU Query (mlg_query_counts_export / _import)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as tdist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, out_dir, mode="dense"):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import synth
    from metalign_b200 import dist as mdist
    from oracle.oracle_c import OracleDB, OracleQuery
    r, w, _ = mdist.init_from_env("gloo")
    assert (r, w) == (rank, world)
    p = synth.params(G=30, n=60, seed=3, len_min=5000, len_max=9000, n_present=6)
    keys = synth.sketch_keys(p)
    nreads = 9001
    a, b = mdist.shard_range(nreads, rank, world)
    bases, nmask = synth.reads_packed(p, a, b - a)                  # this rank's contiguous share
    db = OracleDB(keys, p.G, p.n, 60, (30, 40, 50, 60))
    q = OracleQuery(db, ci_min=2)
    q.push_packed(bases, nmask, None, b - a, p.read_len)
    mdist.check_reducible(2, world)
    if mode == "dense":
        counts = torch.from_numpy(q.export_counts())                   # clamped to ci_min
        assert int(counts.max()) <= 2
        mdist.allreduce_counts(counts)
        q.import_counts(counts.numpy())
    else:
        own = q.export_counts()
        recv, sizes = mdist.allgather_sparse(torch.from_numpy(q.export_sparse()))
        q.merge_sparse(own, [recv[r][:sizes[r]].numpy() for r in range(world) if r != rank])
    res = q.finish()
    np.save(os.path.join(out_dir, "num_%d.npy" % rank), res["num"])
    np.save(os.path.join(out_dir, "ni_%d.npy" % rank), np.array([res["n_intersect"]]))
    tdist.destroy_process_group()


@pytest.mark.parametrize("mode,world", [("dense", 2), ("sparse", 2), ("sparse", 3)])
def test_two_rank_counter_allreduce_equals_single_run(tmp_path, mode, world):
    import synth
    from oracle.oracle_c import OracleDB, OracleQuery
    port = 29500 + (os.getpid() % 400) + {"dense": 0, "sparse": 1}[mode] + 2 * world
    mp.spawn(_worker, args=(world, port, str(tmp_path), mode), nprocs=world, join=True)
    p = synth.params(G=30, n=60, seed=3, len_min=5000, len_max=9000, n_present=6)
    keys = synth.sketch_keys(p)
    bases, nmask = synth.reads_packed(p, 0, 9001)
    db = OracleDB(keys, p.G, p.n, 60, (30, 40, 50, 60))
    q = OracleQuery(db)
    q.push_packed(bases, nmask, None, 9001, p.read_len)
    ref = q.finish()
    assert ref["n_intersect"] > 0
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / ("num_%d.npy" % r)), ref["num"])
        assert int(np.load(tmp_path / ("ni_%d.npy" % r))[0]) == ref["n_intersect"]


def test_shard_range_and_limits():
    from metalign_b200 import dist as mdist
    for n in (0, 1, 7, 100, 101):
        for w in (1, 2, 3, 8):
            spans = [mdist.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    mdist.check_reducible(2, 8)
    with pytest.raises(ValueError):
        mdist.check_reducible(64, 8)
